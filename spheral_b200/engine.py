"""Engine -- thin Python handle over the C ABI (one per GPU / rank).

This is the device boundary of SURVEY.md 8b: everything below it is CUDA, everything above it is the host-side
mirror of Spheral's Physics package interface (spheral_b200/physics.py).  Errors of the C ABI become
RuntimeError (the reference raises VERIFYError -> Python exception, Utilities/DBC.hh:28-45).
"""
import ctypes as C
import numpy as np

from . import _lib as L


class SPHB200Error(RuntimeError):
    pass


def make_options(ndim, **kw):
    o = L.Options()
    o.ndim = ndim
    o.compatibleEnergy, o.evolveTotalEnergy, o.XSPH, o.correctVelocityGradient = 1, 0, 1, 1
    o.epsTensile, o.nTensile, o.nPerh = 0.0, 4.0, 2.01
    o.Qkind, o.Cl, o.Cq, o.eps2, o.negligibleSoundSpeed = L.Q_MG, 1.0, 1.0, 1.0e-2, 1.0e-10
    o.balsara = o.linearInExpansion = o.quadraticInExpansion = 0
    o.etaCritFrac, o.etaFoldFrac = 1.0, 0.2
    o.hEvolution, o.hmin, o.hmax = L.H_SPH, 1.0e-20, 1.0e20
    o.hydro = L.HYDRO_SPH
    o.hminratio = 0.1                        # NodeList default
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def make_step_options(**kw):
    """GammaLawGas(5/3) without pressure limits, the FluidNodeList defaults rhoMin/rhoMax = 1e-10/1e10 and hminratio = 0.1,
    IdealH (NodeList/FluidNodeList.hh, NodeList/NodeList.hh, SmoothingScale/SmoothingScaleBase.hh)."""
    so = L.StepOptions()
    so.eos.gamma = 5.0/3.0
    so.eos.minimumPressure, so.eos.maximumPressure, so.eos.externalPressure, so.eos.minPressureType = -1.0e200, 1.0e200, 0.0, 0
    so.rhoMin, so.rhoMax, so.hminratio, so.HEvolution = 1.0e-10, 1.0e10, 0.1, L.HEVOLUTION_IDEALH
    for k, v in kw.items():
        if hasattr(so.eos, k):
            setattr(so.eos, k, v)
        elif hasattr(so, k):
            setattr(so, k, v)
        else:
            raise KeyError(k)
    return so


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    def __init__(self, ndim, device=0, options=None, **kw):
        self._lib = L.lib()
        self.ndim = ndim
        self.device = device
        self.options = options if options is not None else make_options(ndim, **kw)
        self._h = C.c_void_p()
        if self._lib.sphb200_create(C.byref(self._h), device, C.byref(self.options)) != 0:
            raise SPHB200Error(self._lib.sphb200_last_error(None).decode())
        self.nInternal = self.nGhost = 0
        self.npairs = 0

    # -- plumbing ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise SPHB200Error(self._lib.sphb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.sphb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n(self):
        return self.nInternal + self.nGhost

    def set_options(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.options, k):
                raise KeyError(k)
            setattr(self.options, k, v)
        self._check(self._lib.sphb200_set_options(self._h, C.byref(self.options)))

    def sync(self):
        self._check(self._lib.sphb200_sync(self._h))

    @property
    def stream(self):
        return self._lib.sphb200_stream(self._h)

    # -- kernel tables ---------------------------------------------------------------------------------------
    def set_kernel_table(self, table, which=L.TABLE_W):
        """table: spheral_b200.kernel.TableKernel (or anything with the same attributes)."""
        nperh = getattr(table, "nperhVals", None)
        wsum = getattr(table, "wsumVals", None)
        keep = [np.ascontiguousarray(table.Wcoef), np.ascontiguousarray(table.gradWcoef),
                np.ascontiguousarray(table.grad2Wcoef)]
        nN = len(nperh)//2 if nperh is not None else 0
        wN = len(wsum)//2 if wsum is not None else 0
        self._check(self._lib.sphb200_set_kernel_table(
            self._h, which, table.kernelExtent, table.xmin, table.xstep, table.n1,
            _dp(keep[0]), _dp(keep[1]), _dp(keep[2]),
            nN, table.nperhRange[0] if nN else 0.0, table.nperhRange[1] if nN else 0.0,
            _dp(np.ascontiguousarray(nperh)) if nN else None,
            wN, table.wsumRange[0] if wN else 0.0, table.wsumRange[1] if wN else 0.0,
            _dp(np.ascontiguousarray(wsum)) if wN else None))

    # -- nodes and state ---------------------------------------------------------------------------------------
    def set_nodes(self, nInternal, nGhost=0):
        self._check(self._lib.sphb200_set_nodes(self._h, nInternal, nGhost))
        self.nInternal, self.nGhost = nInternal, nGhost

    def upload_state(self, _keep_connectivity=False, **fields):
        """fields: name -> array in the reference AoS layout; names from _lib.STATE_FIELDS.  _keep_connectivity: the pair lists
        stay valid although positions / H are refreshed (a mid-step evaluation on the step-start ConnectivityMap)."""
        hs = L.HostState()
        mask = 0
        keep = []
        for k, v in fields.items():
            if v is None:
                continue
            if k not in L.STATE_BITS:
                raise KeyError(k)
            a = np.ascontiguousarray(v, dtype=np.float64)
            if a.size != self.n*L.state_width(self.ndim, k):
                raise ValueError("field %s has %d values, expected %d" % (k, a.size, self.n*L.state_width(self.ndim, k)))
            keep.append(a)
            setattr(hs, k, _dp(a))
            mask |= L.STATE_BITS[k]
        fn = self._lib.sphb200_upload_state_values if _keep_connectivity else self._lib.sphb200_upload_state
        self._check(fn(self._h, mask, C.byref(hs)))
        self.sync()          # the numpy temporaries must outlive the async copies

    def upload_state_pinned(self, mask, hs):
        """Asynchronous variant used by bench.py: hs is a prepared HostState over pinned buffers."""
        self._check(self._lib.sphb200_upload_state(self._h, mask, C.byref(hs)))

    def download_state(self, *names):
        arr = (C.POINTER(C.c_double)*len(L.STATE_FIELDS))()
        out = {}
        mask = 0
        for k in names:
            w = L.state_width(self.ndim, k)
            out[k] = np.zeros((self.n, w) if w > 1 else self.n)
            arr[L.STATE_FIELDS.index(k)] = _dp(out[k])
            mask |= L.STATE_BITS[k]
        self._check(self._lib.sphb200_download_state(self._h, mask, arr))
        return out

    # -- connectivity ----------------------------------------------------------------------------------------------
    def build_pairs(self):
        np_ = C.c_size_t()
        self._check(self._lib.sphb200_build_pairs(self._h, C.byref(np_)))
        self.npairs = np_.value
        return self.npairs

    def connectivity_valid(self):
        return bool(self._lib.sphb200_connectivity_valid(self._h))

    def state_fields_present(self):
        """Mask (_lib.STATE_BITS) of the state fields holding values on the device."""
        return int(self._lib.sphb200_state_fields_present(self._h))

    def download_pairs(self):
        pi = np.zeros(max(self.npairs, 1), dtype=np.uint32)
        pj = np.zeros(max(self.npairs, 1), dtype=np.uint32)
        self._check(self._lib.sphb200_download_pairs(self._h, pi.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     pj.ctypes.data_as(C.POINTER(C.c_uint32)), self.npairs))
        return pi[:self.npairs], pj[:self.npairs]

    def download_neighbor_counts(self):
        c = np.zeros(max(self.nInternal, 1), dtype=np.uint32)
        self._check(self._lib.sphb200_download_neighbor_counts(self._h, c.ctypes.data_as(C.POINTER(C.c_uint32))))
        return c[:self.nInternal]

    # -- derivatives ---------------------------------------------------------------------------------------------------
    def evaluate_derivatives(self, time=0.0, dt=1.0):
        self._check(self._lib.sphb200_evaluate_derivatives(self._h, time, dt))

    def download_derivs(self, *names):
        names = names or L.DERIV_FIELDS
        hd = L.HostDerivs()
        out = {}
        mask = 0
        for k in names:
            w = L.deriv_width(self.ndim, k)
            out[k] = np.zeros((self.n, w) if w > 1 else self.n)
            setattr(hd, k, _dp(out[k]))
            mask |= L.DERIV_BITS[k]
        self._check(self._lib.sphb200_download_derivs(self._h, mask, C.byref(hd)))
        return out

    def evaluate_derivatives_to_host(self, *names, time=0.0, dt=1.0):
        """evaluateDerivatives with the selected derivative fields delivered to host arrays, the download of one chunk of the host
        index range overlapped with the pair loop of the next (sphb200_evaluate_derivatives_to_host)."""
        names = names or L.DERIV_FIELDS
        hd = L.HostDerivs()
        out = {}
        mask = 0
        for k in names:
            w = L.deriv_width(self.ndim, k)
            out[k] = np.zeros((self.n, w) if w > 1 else self.n)
            setattr(hd, k, _dp(out[k]))
            mask |= L.DERIV_BITS[k]
        self._check(self._lib.sphb200_evaluate_derivatives_to_host(self._h, time, dt, mask, C.byref(hd)))
        return out

    def upload_derivs(self, **fields):
        """Restart (SPHBase::restoreState): node-wise derivative fields back onto the device; names from _lib.DERIV_FIELDS."""
        hd = L.HostDerivs()
        mask, keep = 0, []
        for k, v in fields.items():
            if k not in L.DERIV_BITS:
                raise KeyError(k)
            a = np.ascontiguousarray(v, dtype=np.float64)
            if a.size != self.n*L.deriv_width(self.ndim, k):
                raise ValueError("field %s has %d values, expected %d" % (k, a.size, self.n*L.deriv_width(self.ndim, k)))
            keep.append(a)
            setattr(hd, k, _dp(a))
            mask |= L.DERIV_BITS[k]
        self._check(self._lib.sphb200_upload_derivs(self._h, mask, C.byref(hd)))

    def download_pair_accelerations(self):
        out = np.zeros((max(self.npairs, 1), self.ndim))
        self._check(self._lib.sphb200_download_pair_accelerations(self._h, _dp(out), out.size))
        return out[:self.npairs]

    def copy_DvDx_to_Q(self):
        self._check(self._lib.sphb200_copy_DvDx_to_Q(self._h))

    def update_energy_compatible(self, multiplier):
        self._check(self._lib.sphb200_update_energy_compatible(self._h, multiplier))

    # -- CRKSPH (contexts created with hydro=HYDRO_CRKSPH) -----------------------------------------------------------
    def crk_compute_volume(self):
        """computeRKSumVolume -> internal entries of 'volume' on the device."""
        self._check(self._lib.sphb200_crk_compute_volume(self._h))

    def crk_compute_corrections(self):
        """RKUtilities::computeCorrections (LinearOrder) -> internal entries of 'rkCorrections' on the device."""
        self._check(self._lib.sphb200_crk_compute_corrections(self._h))

    def crk_sum_mass_density(self, rhoMin=0.0, rhoMax=1.0e300):
        """computeCRKSPHSumMassDensity -> internal entries of 'massDensity' on the device."""
        self._check(self._lib.sphb200_crk_sum_mass_density(self._h, rhoMin, rhoMax))

    # -- per-step callers, device-resident ------------------------------------------------------------
    def sum_mass_density(self):
        """computeSPHSumMassDensity -> internal entries of 'massDensity' on the device."""
        self._check(self._lib.sphb200_sum_mass_density(self._h))

    def compute_omega_gradh(self):
        """computeSPHOmegaGradhCorrection -> internal entries of 'omegaGradh' on the device."""
        self._check(self._lib.sphb200_compute_omega_gradh(self._h))

    def update_eos_gamma_law(self, eos):
        """PressurePolicy / SoundSpeedPolicy with GammaLawGas over every node; eos: _lib.GammaLaw."""
        self._check(self._lib.sphb200_update_eos_gamma_law(self._h, C.byref(eos)))

    def state_copy(self):
        self._check(self._lib.sphb200_state_copy(self._h))

    def state_assign(self):
        self._check(self._lib.sphb200_state_assign(self._h))

    def state_update(self, step_options, multiplier, timeAdvanceOnly=False):
        """State::update(derivs, multiplier, t, dt) for the hydro + smoothing-scale policies, then P and cs."""
        self._check(self._lib.sphb200_state_update(self._h, C.byref(step_options), multiplier, int(bool(timeAdvanceOnly))))

    def iterate_ideal_h(self, first_sweep, tolerance):
        """One H <- 'new H' sweep of iterateIdealH over the nodes not yet converged; returns maxDeltaH."""
        d = C.c_double()
        self._check(self._lib.sphb200_iterate_ideal_h(self._h, int(bool(first_sweep)), tolerance, C.byref(d)))
        return d.value

    def compute_dt(self, cfl=0.25, useVelocityMagnitudeForDt=False):
        """GenericHydro::dt -> (dt, reason string, limiting node)."""
        dt, reason, node = C.c_double(), C.c_int(), C.c_uint32()
        self._check(self._lib.sphb200_compute_dt(self._h, cfl, int(bool(useVelocityMagnitudeForDt)), C.byref(dt), C.byref(reason),
                                                 C.byref(node)))
        return dt.value, (L.DT_REASONS[reason.value] if reason.value >= 0 else ""), node.value

    # -- reflecting planes on the device -----------------------------------------------------------------
    def reflect_configure(self, planes):
        """planes: list of (point, inward normal)."""
        pts = np.ascontiguousarray([p for p, _ in planes], dtype=np.float64).reshape(-1)
        nrm = np.ascontiguousarray([n for _, n in planes], dtype=np.float64).reshape(-1)
        self._check(self._lib.sphb200_reflect_configure(self._h, len(planes), _dp(pts) if len(planes) else None,
                                                        _dp(nrm) if len(planes) else None))

    def boundary_configure(self, boundaries):
        """boundaries: list of ("reflecting", (point, normal)) or ("periodic", (point1, normal1), (point2, normal2)); a periodic
        boundary expands to its two planar boundaries (plane1 -> plane2, plane2 -> plane1; PeriodicBoundary.cc:60-62)."""
        kinds, ep, en, xp, xn = [], [], [], [], []
        for b in boundaries:
            if b[0] == "reflecting":
                kinds.append(0); ep.append(b[1][0]); en.append(b[1][1]); xp.append(b[1][0]); xn.append(b[1][1])
            elif b[0] == "periodic":
                (p1, n1), (p2, n2) = b[1], b[2]
                kinds += [1, 1]; ep += [p1, p2]; en += [n1, n2]; xp += [p2, p1]; xn += [n2, n1]
            else:
                raise ValueError("unknown boundary kind %r" % (b[0],))
        arr = lambda v: np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
        k = (C.c_int*max(len(kinds), 1))(*kinds)
        a = [arr(v) for v in (ep, en, xp, xn)]
        self._keep_bnd = a
        self._check(self._lib.sphb200_boundary_configure(self._h, len(kinds), k, *[_dp(v) if len(kinds) else None for v in a]))

    def reflect_set_ghost_nodes(self):
        """PlanarBoundary::setGhostNodes for every plane; returns (and records) the new ghost count."""
        ng = C.c_size_t()
        self._check(self._lib.sphb200_reflect_set_ghost_nodes(self._h, C.byref(ng)))
        self.nGhost = ng.value
        return self.nGhost

    def reflect_apply_ghosts(self, mask=None):
        """ReflectingBoundary::applyGhostBoundary for the masked state fields (default: every field on the device)."""
        self._check(self._lib.sphb200_reflect_apply_ghosts(self._h, 0xFFFFFFFF if mask is None else mask))

    def reflect_finalize_derivatives(self):
        """SPHBase::finalizeDerivatives: ghost values of DvDt and DepsDt (needed by the compatible energy update)."""
        self._check(self._lib.sphb200_reflect_finalize_derivatives(self._h))

    def reflect_enforce(self, count=False):
        nv = C.c_size_t()
        self._check(self._lib.sphb200_reflect_enforce(self._h, C.byref(nv) if count else None))
        return nv.value

    # -- halo ----------------------------------------------------------------------------------------------------------
    def halo_bytes_per_node(self, mask):
        return self._lib.sphb200_halo_bytes_per_node(self._h, mask)

    def halo_pack(self, mask, send_nodes_ptr, count, staging_ptr):
        self._check(self._lib.sphb200_halo_pack(self._h, mask, send_nodes_ptr, count, staging_ptr))

    def halo_unpack(self, mask, first_ghost, count, staging_ptr):
        self._check(self._lib.sphb200_halo_unpack(self._h, mask, first_ghost, count, staging_ptr))

    def halo_unpack_values(self, mask, first_ghost, count, staging_ptr):
        """Ghost-value refresh that keeps the connectivity (applyGhostBoundaries between the stages of a step)."""
        self._check(self._lib.sphb200_halo_unpack_values(self._h, mask, first_ghost, count, staging_ptr))

    def halo_pack_derivs(self, send_nodes_ptr, count, staging_ptr):
        self._check(self._lib.sphb200_halo_pack_derivs(self._h, send_nodes_ptr, count, staging_ptr))

    def halo_unpack_derivs(self, first_ghost, count, staging_ptr):
        self._check(self._lib.sphb200_halo_unpack_derivs(self._h, first_ghost, count, staging_ptr))

    def node_bounds(self, count=None):
        """(lo[ndim], hi[ndim], maxExtent[ndim]) of nodes [0,count) -- default: the internal nodes."""
        lo, hi, ext = np.zeros(3), np.zeros(3), np.zeros(3)
        self._check(self._lib.sphb200_node_bounds(self._h, self.nInternal if count is None else count, _dp(lo), _dp(hi), _dp(ext)))
        return lo[:self.ndim], hi[:self.ndim], ext[:self.ndim]

    def halo_select(self, axis, lo, hi, width, send_low_ptr, send_high_ptr, cap, count=None):
        """Device-side send-node selection of a slab decomposition; returns (nLow, nHigh)."""
        nl, nh = C.c_size_t(), C.c_size_t()
        self._check(self._lib.sphb200_halo_select(self._h, axis, self.nInternal if count is None else count, lo, hi, width,
                                                  send_low_ptr, C.byref(nl), send_high_ptr, C.byref(nh), cap))
        return nl.value, nh.value

    def node_bounds_device(self, count, bounds_ptr):
        """Stream-ordered, no host sync: {lo[3], hi[3], maxExtent[3]} of nodes [0,count) into a 9-double device buffer."""
        self._check(self._lib.sphb200_node_bounds_device(self._h, count, bounds_ptr))

    def halo_select_device(self, axis, lo, hi, max_extent_ptr, send_low_ptr, send_high_ptr, counts_ptr, cap, count=None):
        """Stream-ordered, no host sync: the halo width is read from the device; {nLow, nHigh} (int64) land in counts_ptr."""
        self._check(self._lib.sphb200_halo_select_device(self._h, axis, self.nInternal if count is None else count, lo, hi,
                                                         max_extent_ptr, send_low_ptr, send_high_ptr, counts_ptr, cap))

    # -- instrumentation -----------------------------------------------------------------------------------------------
    def stats(self):
        s = L.Stats()
        self._check(self._lib.sphb200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in L.Stats._fields_}

    def measure_fp64_peak(self):
        t = C.c_double()
        self._check(self._lib.sphb200_measure_fp64_peak(self._h, C.byref(t)))
        return t.value
