"""Spatial domain decomposition and ghost-node (halo) exchange for the SPH hot path on 1-8 GPUs of one box.

Replaces, for this path, the MPI ghost exchange of src/Distributed:
    DistributedBoundary::setAllGhostNodes / buildReceiveAndGhostNodes   (Distributed/DistributedBoundary.cc:565-753, 1256-1353)
    TreeDistributedBoundary::setAllGhostNodes                            (Distributed/TreeDistributedBoundary.cc:62-91, 150-295)
with one process per GPU, `torch.distributed` point-to-point operations (NCCL send/recv over NVLink on GPUs; gloo in the
CPU tests) and device-side pack / unpack kernels reached through the C ABI (sphb200_halo_select / _pack / _unpack).

Decomposition: 1-D slabs along one axis -- two peers per rank.  A node is sent to a neighbouring slab when it
lies within `width` of the shared face, where `width` is the largest per-axis kernel extent kext*sqrt((H^-2)_aa) over ALL
ranks (an all-reduce MAX), so that both the gather (H_i) and the scatter (H_j) side of the reference's pair predicate
(ConnectivityMap.cc:916-931) are covered; the ghost set is a superset of what the reference would send, never a subset.

The derivative evaluation itself needs no communication: ghosts are filled beforehand and derivatives are only defined on
internal nodes.  The exchange is split in two so that only positions and H sit on the critical path:
    phase A  position, H            -> needed by the neighbour build (K1 + K2)
    phase B  every other state field -> needed by the pair loop (K3); in flight while K1 + K2 run.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L

PHASE_A = ("position", "H")
PHASE_B = ("velocity", "mass", "massDensity", "specificThermalEnergy", "pressure", "soundSpeed", "omegaGradh")
# State fields owned by the artificial viscosity (ArtificialViscosityHandle.cc:105-119): they travel with the halo from the moment they
# hold values on the device -- LimitedMonaghanGingold and the Balsara switch read the neighbour's velocity gradient, ghosts included
OPTIONAL_B = ("DvDxQ", "fCl", "fCq")


def slab_edges(xmin, xmax, world):
    """Equal-width slab faces along the decomposition axis: world+1 values."""
    return np.linspace(float(xmin), float(xmax), world + 1)


def kernel_extent_axis(H, ndim, kext, axis):
    """kext*sqrt((H^-2)_aa) per node (Neighbor::HExtent, NeighborInline.hh:52-64) -- numpy reference of what
    sphb200_node_bounds reduces on the device."""
    H = np.asarray(H, dtype=np.float64)
    n = H.shape[0]
    F = np.zeros((n, ndim, ndim))
    if ndim == 3:
        idx = ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))
    else:
        idx = ((0, 0), (0, 1), (1, 1))
    for k, (r, c) in enumerate(idx):
        F[:, r, c] = H[:, k]
        F[:, c, r] = H[:, k]
    Fi = np.linalg.inv(F)
    return kext*np.sqrt((Fi[:, axis, :]**2).sum(axis=1))


def select_halo_numpy(pos, count, axis, lo, hi, width):
    """Host reference of sphb200_halo_select: ascending indices of nodes [0,count) to send to the lower / upper slab."""
    x = np.asarray(pos)[:count, axis]
    return (np.nonzero(x < lo + width)[0].astype(np.uint32), np.nonzero(x >= hi - width)[0].astype(np.uint32))


def send_list_needs(allc, world):
    """From the all-gathered table {nLow, nHigh, capacity} of every rank: the send-list length each rank really needs (the low list
    of the first slab and the high list of the last one are never sent) and whether every list fitted its rank's capacity.  Every
    rank evaluates this on the same table, so all of them take the same decision to redo the selection or to go on -- a rank-local
    test lets one slab repeat the all-gather while its neighbour is already in send/recv."""
    need = [max(int(allc[r][0]) if r > 0 else 0, int(allc[r][1]) if r < world - 1 else 0) for r in range(world)]
    return need, all(need[r] <= int(allc[r][2]) for r in range(world))


class SlabHalo:
    """Point-to-point plumbing between a slab and its two neighbours.  Works on torch tensors, so the same code runs over
    NCCL (CUDA tensors) and gloo (CPU tensors, used by tests/test_distributed_gloo.py)."""

    def __init__(self, rank=None, world=None, group=None):
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.group = group
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank < self.world - 1 else None

    def allreduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def exchange_counts(self, nLow, nHigh, device):
        """Tell each neighbour how many nodes it will receive; returns (nFromLower, nFromUpper)."""
        if self.world == 1:
            return 0, 0
        mine = torch.tensor([nLow, nHigh], dtype=torch.int64, device=device)
        parts = [torch.empty(2, dtype=torch.int64, device=device) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        allc = torch.stack(parts).cpu().numpy()
        nFromLower = int(allc[self.lower, 1]) if self.lower is not None else 0     # the lower slab's "high" list is ours
        nFromUpper = int(allc[self.upper, 0]) if self.upper is not None else 0
        return nFromLower, nFromUpper

    def start(self, sendLow, sendHigh, recvLow, recvHigh):
        """Post the sends and receives of one phase; returns the work handles (wait() orders the current stream after them)."""
        ops = []
        if self.lower is not None:
            if sendLow is not None and sendLow.numel():
                ops.append(dist.P2POp(dist.isend, sendLow, self.lower, self.group))
            if recvLow is not None and recvLow.numel():
                ops.append(dist.P2POp(dist.irecv, recvLow, self.lower, self.group))
        if self.upper is not None:
            if sendHigh is not None and sendHigh.numel():
                ops.append(dist.P2POp(dist.isend, sendHigh, self.upper, self.group))
            if recvHigh is not None and recvHigh.numel():
                ops.append(dist.P2POp(dist.irecv, recvHigh, self.upper, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(works):
        for w in works:
            w.wait()


def ordered_fields(names):
    """Fields in the order sphb200_halo_pack lays them out: ascending mask bit."""
    return [k for k in L.STATE_FIELDS if k in names]


def pack_fields_numpy(state, names, idx, ndim):
    """Host reference of sphb200_halo_pack: field-major staging, each field len(idx)*width doubles."""
    return np.concatenate([np.asarray(state[k], dtype=np.float64).reshape(-1, L.state_width(ndim, k))[idx].ravel()
                           for k in ordered_fields(names)]) if len(idx) else np.zeros(0)


def unpack_fields_numpy(buf, names, count, ndim):
    """Host reference of sphb200_halo_unpack: staging -> {field: (count, width) array}."""
    out, off = {}, 0
    for k in ordered_fields(names):
        w = L.state_width(ndim, k)
        out[k] = np.asarray(buf[off:off + count*w]).reshape(count, w) if w > 1 else np.asarray(buf[off:off + count])
        off += count*w
    return out


def field_mask(names):
    m = 0
    for k in names:
        m |= L.STATE_BITS[k]
    return m


class DistributedSPH:
    """One slab of a domain-decomposed problem on one GPU: an Engine plus the ghost exchange with the two neighbouring
    slabs.  `step_connectivity_and_derivatives()` is what Integrator::setGhostNodes + evaluateDerivatives do per stage
    (Integrator.cc:372-445, 217-229): ghost selection, exchange, neighbour build, derivative evaluation."""

    def __init__(self, engine, axis, lo, hi, halo=None, extra_fields=(), two_phase=None):
        self.e = engine
        # default: split exchange (positions + H first, the rest in flight during the neighbour build) from 2 M internal nodes per rank on,
        # one batch below -- the split costs two more pack / unpack launches, a second NCCL batch and a partial re-pack of the rows, which a
        # small slab does not win back (2 ranks x 1 M: 4.31 ms split, 4.20 ms single batch; 2 ranks x 4 M: 16.61 / 16.53 ms).
        # SPHB200_HALO_TWO_PHASE=0 / 1 forces one or the other (A/B measurements)
        env = os.environ.get("SPHB200_HALO_TWO_PHASE")
        if two_phase is not None:
            self.two_phase = bool(two_phase)
        elif env is not None:
            self.two_phase = env != "0"
        else:
            self.two_phase = None                    # decided below from the smallest slab, identically on every rank
        self.axis, self.lo, self.hi = axis, float(lo), float(hi)
        self.halo = halo if halo is not None else SlabHalo()
        self.dev = torch.device("cuda", engine_device(engine))
        self.stream = torch.cuda.ExternalStream(engine.stream, device=self.dev)
        if self.two_phase is None:
            t = torch.tensor([engine.nInternal], dtype=torch.int64, device=self.dev)
            if self.halo.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.halo.group)
            self.two_phase = int(t.item()) >= (1 << 21)
        self._timing, self.cpu_ms, self._t = os.environ.get("SPHB200_HALO_TIMING", "0") != "0", {}, 0.0
        self.maskA = field_mask(PHASE_A)
        # extra_fields: state fields that only exist once a package has computed them (CRKSPH: "volume", "rkCorrections").  They join
        # the exchange from the moment the integrator reports them ready (mark_ready); the staging is sized for all of them.
        self.extra = tuple(extra_fields)
        self.ready = set()
        self._maskB_base = field_mask(PHASE_B)
        self._maskB_opt = field_mask(OPTIONAL_B)
        self.bytesA = engine.halo_bytes_per_node(self.maskA)
        self.bytesB = engine.halo_bytes_per_node(field_mask(PHASE_B + OPTIONAL_B + self.extra))
        self._cap = 0
        self.nInternal = engine.nInternal
        # ghosts of the boundaries that come before the slab halo in the boundary list (reflecting / periodic planes generated on
        # the device by sphb200_reflect_set_ghost_nodes): they sit right after the internal nodes, take part in the send-node
        # selection like internal nodes (DistributedBoundary exchanges the ghosts of the other boundaries too, so that a node near
        # a slab face AND a plane sees the mirror images owned by the neighbouring slab) and the halo lands after them
        self.nBoundaryGhost = 0
        self.nFromLower = self.nFromUpper = 0
        self.width = 0.0
        self.last = {}

    @property
    def maskB(self):
        """Every field but positions and H that travels: the hydro's state, the artificial viscosity's fields and the extra fields
        declared ready -- each only if it holds values on the device (CRKSPH, for one, has no grad-h correction field; every rank runs
        the same call sequence, so all ranks agree on the set)."""
        m = self._maskB_base | self._maskB_opt | field_mask(tuple(x for x in self.extra if x in self.ready))
        return m & self.e.state_fields_present()

    def mark_ready(self, *names):
        """A package has computed these extra fields on the internal nodes: from now on they travel with the halo."""
        for k in names:
            if k in self.extra:
                self.ready.add(k)

    def _ensure(self, cap, keep_lists=False):
        """Capacity (in nodes) of the send lists and of the four staging buffers.  keep_lists: the send lists have just been
        selected and must survive the growth (the receive side of a slab may need more room than its own send side)."""
        if cap <= self._cap:
            return
        cap = int(cap*1.25) + 1024
        dev = self.dev
        old = (self.idxLow, self.idxHigh) if (keep_lists and self._cap) else None
        self.idxLow = torch.empty(cap, dtype=torch.int32, device=dev)
        self.idxHigh = torch.empty(cap, dtype=torch.int32, device=dev)
        if old is not None:
            self.idxLow[:old[0].numel()].copy_(old[0])
            self.idxHigh[:old[1].numel()].copy_(old[1])
        mk = lambda nbytes: torch.empty(cap*nbytes//8, dtype=torch.float64, device=dev)
        self.sLowA, self.sHighA, self.rLowA, self.rHighA = mk(self.bytesA), mk(self.bytesA), mk(self.bytesA), mk(self.bytesA)
        nb = self.bytesA + self.bytesB            # the single-batch exchange stages every field in the B buffers
        self.sLowB, self.sHighB, self.rLowB, self.rHighB = mk(nb), mk(nb), mk(nb), mk(nb)
        self._cap = cap

    def refresh_ghosts_and_build(self):
        return self.refresh_ghosts(build=True)

    def _tick(self, name):
        """SPHB200_HALO_TIMING=1: host-side time between the marks of refresh_ghosts, accumulated in self.cpu_ms (diagnostic: on small
        slabs the host enqueue path, not the GPU, bounds the ghost refresh)."""
        if not self._timing:
            return
        import time
        now = time.perf_counter()
        if name is not None:
            self.cpu_ms[name] = self.cpu_ms.get(name, 0.0) + (now - self._t)*1e3
        self._t = now

    def refresh_ghosts(self, build=True, boundary_ghosts=0):
        """Ghost selection + exchange (+ neighbour build).  Returns the number of node pairs of this slab (None without the build).
        `boundary_ghosts`: number of plane ghosts already generated behind the internal nodes (see nBoundaryGhost).

        ONE host round trip before the neighbour build: the bounds reduction, the all-reduce(MAX) of the halo width, the
        send-node selection and the all-gather of the counts are chained on the device (sphb200_node_bounds_device /
        sphb200_halo_select_device); only the gathered counts come back to the host (NCCL needs the message sizes there).
        Default (two_phase): positions + H travel first and the neighbour build (K1 + K2) starts as soon as they have landed;
        every other field is in flight on the NCCL stream meanwhile and lands before the pair loop, where only the rows of the
        ghosts it touched are packed again (sphb200_halo_unpack -> k_pack range mode; round 1 re-packed all rows, which cost more
        than the overlap saved).  two_phase=False sends everything in one batch before the build."""
        e, h = self.e, self.halo
        nInt = self.nInternal
        if nInt == 0:
            raise RuntimeError("DistributedSPH: a slab without internal nodes is not supported")
        tick = self._tick
        tick(None)
        nBG = self.nBoundaryGhost = int(boundary_ghosts)
        nOwn = nInt + nBG                                  # nodes this slab can send: internal + its own plane ghosts
        with torch.cuda.stream(self.stream):
            if self._cap == 0:
                self._ensure(max(1024, nInt//8))
                self._bounds = torch.zeros(9, dtype=torch.float64, device=self.dev)
                self._counts = torch.zeros(3, dtype=torch.int64, device=self.dev)      # {nLow, nHigh, capacity of the send lists}
                self._allc = torch.zeros(3*h.world, dtype=torch.int64, device=self.dev)
            while True:
                # halo width: largest kernel extent along the axis over all ranks (gather AND scatter neighbours are covered)
                e.node_bounds_device(nOwn, self._bounds.data_ptr())
                ext = self._bounds[6:9]
                h.allreduce_max(ext)
                tick("bounds + all-reduce enqueued")
                e.halo_select_device(self.axis, self.lo, self.hi, ext.data_ptr(), self.idxLow.data_ptr(), self.idxHigh.data_ptr(),
                                     self._counts.data_ptr(), self._cap, nOwn)
                if h.world > 1:
                    dist.all_gather_into_tensor(self._allc, self._counts, group=h.group)
                else:
                    self._allc.copy_(self._counts)
                tick("select + all-gather enqueued")
                allc = self._allc.cpu().numpy().reshape(h.world, 3)          # the one host round trip
                tick("counts on the host (sync)")
                nLow, nHigh = int(allc[h.rank, 0]), int(allc[h.rank, 1])
                # Redo the selection if ANY rank truncated a send list it will use.  The decision is taken from the gathered table,
                # identically on every rank: a rank-local test would let one slab repeat the all-gather while its neighbour moves
                # on to the send/recv (a deadlock; seen when plane ghosts made the counts of the end slabs asymmetric).
                need, fits = send_list_needs(allc, h.world)
                if fits:
                    break
                self._ensure(max(need))                                       # send lists were truncated: grow and redo
            if h.lower is None:
                nLow = 0
            if h.upper is None:
                nHigh = 0
            nFL = int(allc[h.lower, 1]) if h.lower is not None else 0         # the lower slab's "high" list is ours
            nFU = int(allc[h.upper, 0]) if h.upper is not None else 0
            self._ensure(max(nLow, nHigh, nFL, nFU), keep_lists=True)
            self.nFromLower, self.nFromUpper = nFL, nFU
            e.set_nodes(nInt, nBG + nFL + nFU)
            maskB = self.maskB                                  # evaluated once per refresh: landing ghosts must not change it mid-way
            wA, wB = self.bytesA//8, e.halo_bytes_per_node(maskB)//8
            if not self.two_phase or not build:
                wAB = wA + wB
                mask = self.maskA | maskB
                e.halo_pack(mask, self.idxLow.data_ptr(), nLow, self.sLowB.data_ptr())
                e.halo_pack(mask, self.idxHigh.data_ptr(), nHigh, self.sHighB.data_ptr())
                tick("packs enqueued")
                works = h.start(self.sLowB[:nLow*wAB], self.sHighB[:nHigh*wAB], self.rLowB[:nFL*wAB], self.rHighB[:nFU*wAB])
                tick("send/recv batch enqueued")
                h.finish(works)
                e.halo_unpack(mask, nOwn, nFL, self.rLowB.data_ptr())
                e.halo_unpack(mask, nOwn + nFL, nFU, self.rHighB.data_ptr())
                tick("wait + unpacks enqueued")
                npairs = e.build_pairs() if build else None
                tick("build_pairs (sync)")
            else:
                # pack both phases, then post A and B; K1+K2 only wait for A
                e.halo_pack(self.maskA, self.idxLow.data_ptr(), nLow, self.sLowA.data_ptr())
                e.halo_pack(self.maskA, self.idxHigh.data_ptr(), nHigh, self.sHighA.data_ptr())
                e.halo_pack(maskB, self.idxLow.data_ptr(), nLow, self.sLowB.data_ptr())
                e.halo_pack(maskB, self.idxHigh.data_ptr(), nHigh, self.sHighB.data_ptr())
                worksA = h.start(self.sLowA[:nLow*wA], self.sHighA[:nHigh*wA], self.rLowA[:nFL*wA], self.rHighA[:nFU*wA])
                worksB = h.start(self.sLowB[:nLow*wB], self.sHighB[:nHigh*wB], self.rLowB[:nFL*wB], self.rHighB[:nFU*wB])
                h.finish(worksA)
                e.halo_unpack(self.maskA, nOwn, nFL, self.rLowA.data_ptr())
                e.halo_unpack(self.maskA, nOwn + nFL, nFU, self.rHighA.data_ptr())
                npairs = e.build_pairs()                      # phase B is in flight on the NCCL stream meanwhile
                h.finish(worksB)
                e.halo_unpack(maskB, nOwn, nFL, self.rLowB.data_ptr())
                e.halo_unpack(maskB, nOwn + nFL, nFU, self.rHighB.data_ptr())
        self.last = dict(nSendLow=nLow, nSendHigh=nHigh, nFromLower=nFL, nFromUpper=nFU, nBoundaryGhost=nBG,
                         h2h_bytes=(nLow + nHigh)*(self.bytesA + 8*wB), two_phase=bool(self.two_phase))
        return npairs

    def apply_ghosts(self, names=None):
        """Integrator::applyGhostBoundaries on the ghost set of the last refresh: the current values of the send nodes travel
        again and land in the same ghost slots; the connectivity is kept (DistributedBoundary::applyGhostBoundary,
        Distributed/DistributedBoundary.cc:210-330)."""
        e, h = self.e, self.halo
        nInt, L_ = self.nInternal + self.nBoundaryGhost, self.last
        nLow, nHigh, nFL, nFU = L_["nSendLow"], L_["nSendHigh"], L_["nFromLower"], L_["nFromUpper"]
        mask = (self.maskA | self.maskB) if names is None else field_mask(names)
        w = e.halo_bytes_per_node(mask)//8
        with torch.cuda.stream(self.stream):
            e.halo_pack(mask, self.idxLow.data_ptr(), nLow, self.sLowB.data_ptr())
            e.halo_pack(mask, self.idxHigh.data_ptr(), nHigh, self.sHighB.data_ptr())
            h.finish(h.start(self.sLowB[:nLow*w], self.sHighB[:nHigh*w], self.rLowB[:nFL*w], self.rHighB[:nFU*w]))
            e.halo_unpack_values(mask, nInt, nFL, self.rLowB.data_ptr())
            e.halo_unpack_values(mask, nInt + nFL, nFU, self.rHighB.data_ptr())

    def finalize_derivatives(self):
        """SPHBase::finalizeDerivatives (SPHBase.cc:502-519) across the slab faces: the ghost entries of DvDt and DepsDt become
        their owners' values; SpecificThermalEnergyPolicy reads them for the ghost end of an internal-ghost pair."""
        e, h = self.e, self.halo
        nInt, L_ = self.nInternal + self.nBoundaryGhost, self.last
        nLow, nHigh, nFL, nFU = L_["nSendLow"], L_["nSendHigh"], L_["nFromLower"], L_["nFromUpper"]
        w = e.ndim + 1
        with torch.cuda.stream(self.stream):
            e.halo_pack_derivs(self.idxLow.data_ptr(), nLow, self.sLowB.data_ptr())
            e.halo_pack_derivs(self.idxHigh.data_ptr(), nHigh, self.sHighB.data_ptr())
            h.finish(h.start(self.sLowB[:nLow*w], self.sHighB[:nHigh*w], self.rLowB[:nFL*w], self.rHighB[:nFU*w]))
            e.halo_unpack_derivs(nInt, nFL, self.rLowB.data_ptr())
            e.halo_unpack_derivs(nInt + nFL, nFU, self.rHighB.data_ptr())

    def allreduce_min(self, x):
        """Integrator::selectDt's allReduce(dt, MIN) (Integrator.cc:150-154)."""
        if self.halo.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.halo.group)
        return float(t.item())

    def info(self):
        """Counts of the last refresh plus the halo width (read back from the device here, outside the hot path)."""
        if self._cap:
            self.width = float(self._bounds[6 + self.axis].item())*(1.0 + 1.0e-9)
        return dict(self.last, width=self.width)

    def step_connectivity_and_derivatives(self, time=0.0, dt=1.0):
        npairs = self.refresh_ghosts_and_build()
        self.e.evaluate_derivatives(time, dt)
        return npairs


def engine_device(engine):
    return int(getattr(engine, "device", 0))
