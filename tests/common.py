"""Shared problem builders and comparison helpers for the parity tests (tests only; may use the oracle)."""
import math
import numpy as np

from spheral_b200 import nodegen as ng

STATE_KEYS = ("position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy", "pressure", "soundSpeed",
              "omegaGradh", "DvDxQ", "fCl", "fCq")
# product name -> oracle name
ORC_STATE = dict(position="pos", velocity="vel", H="H", mass="mass", massDensity="rho", pressure="P", soundSpeed="cs",
                 omegaGradh="omega", DvDxQ="DvDxQ", fCl="fCl", fCq="fCq")


def smooth_velocity(pos, amp=1.0, seed=7):
    rng = np.random.default_rng(seed)
    nd = pos.shape[1]
    k = rng.uniform(1.0, 4.0, size=(nd, nd))
    ph = rng.uniform(0, 2*math.pi, size=nd)
    v = np.stack([amp*np.sin((pos*k[a]).sum(axis=1) + ph[a]) for a in range(nd)], axis=1)
    return v + 0.05*amp*rng.standard_normal(v.shape)


def make_problem(ndim=3, n=10, nPerh=1.51, kind="lattice", seed=11, negP=False, gamma=5.0/3.0, ghosts=False, kext=2.0):
    """Returns (state dict with product names, nInternal, nGhost)."""
    rng = np.random.default_rng(seed)
    if kind == "lattice":
        pos, mass, H, d = ng.lattice(ndim, n, nPerh=nPerh)
        pos = ng.jitter(pos, 0.2, d, seed=seed)
    elif kind == "aniso":
        N = n**ndim
        pos, H = ng.random_anisotropic(ndim, N, [[0, 1]]*ndim, nPerh=nPerh, seed=seed)
        mass = np.full(N, 1.0/N)
    else:
        raise ValueError(kind)
    N = pos.shape[0]
    rho = 1.0 + 0.3*np.sin(3.0*pos[:, 0])*np.cos(2.0*pos[:, 1]) + 0.02*rng.standard_normal(N)
    mass = mass*rho if kind == "lattice" else mass*(1.0 + 0.1*rng.standard_normal(N))
    eps = 1.0 + 0.5*np.cos(2.5*pos[:, 1]) + 0.05*rng.standard_normal(N)
    vel = smooth_velocity(pos, seed=seed + 1)
    P, cs = ng.gamma_law(rho, eps, gamma)
    if negP:
        P = P - 1.2*np.abs(P).mean()*(rng.uniform(size=N) < 0.3)
    omega = 1.0 + 0.1*rng.standard_normal(N)
    st = dict(position=pos, velocity=vel, H=H, mass=mass, massDensity=rho, specificThermalEnergy=eps, pressure=P,
              soundSpeed=cs, omegaGradh=omega)
    nInt, nGhost = N, 0
    if ghosts:
        planes = [(np.zeros(ndim), np.eye(ndim)[a]) for a in range(ndim)]
        f = dict(pos=pos, H=H, vel=vel, mass=mass, rho=rho, eps=eps, P=P, cs=cs, omega=omega)
        out, ctl, n0 = ng.reflect_ghosts(ndim, f, planes, kext)
        st = dict(position=out["pos"], velocity=out["vel"], H=out["H"], mass=out["mass"], massDensity=out["rho"],
                  specificThermalEnergy=out["eps"], pressure=out["P"], soundSpeed=out["cs"], omegaGradh=out["omega"])
        nInt, nGhost = n0, out["pos"].shape[0] - n0
    return {k: np.ascontiguousarray(v) for k, v in st.items()}, nInt, nGhost


def add_q_fields(st, ndim, seed=3):
    rng = np.random.default_rng(seed)
    N = st["position"].shape[0]
    st = dict(st)
    st["DvDxQ"] = np.ascontiguousarray(0.8*rng.standard_normal((N, ndim*ndim)))
    st["fCl"] = 1.0 + 0.2*rng.uniform(size=N)
    st["fCq"] = 1.0 + 0.3*rng.uniform(size=N)
    return st


def to_oracle_state(st):
    return {ORC_STATE[k]: v for k, v in st.items() if k in ORC_STATE}


def oracle_table(orc, tk):
    """Wrap the PRODUCT's table so that GPU and oracle evaluate one and the same table (SURVEY 7 hard part 2)."""
    return orc.TableKernel.from_arrays(tk.ndim, tk.kernelExtent, tk.xmin, tk.xstep, tk.n1, tk.Wcoef, tk.gradWcoef,
                                       tk.grad2Wcoef, nperh=(tk.nperhVals, tk.nperhRange[0], tk.nperhRange[1]),
                                       wsum=(tk.wsumVals, tk.wsumRange[0], tk.wsumRange[1]))


def opts_pair(orc, eng_mod, ndim, **kw):
    """Identical option sets for the oracle and the product."""
    return orc.default_options(ndim, **kw), eng_mod.make_options(ndim, **kw)


def field_err(a, b, nInt, floor):
    """SURVEY 8c metric: max-norm error over internal nodes, scaled by max(|ref|_inf, floor)."""
    a = np.asarray(a)[:nInt]
    b = np.asarray(b)[:nInt]
    scale = max(float(np.abs(b).max()) if b.size else 0.0, floor)
    return float(np.abs(a - b).max())/scale if b.size else 0.0


def physical_floors(st, nInt, ndim):
    """Per-field physical scales used as floors so that symmetric (near-zero) fields do not produce 0/0."""
    h = 1.0/st["H"][:nInt, 0].mean()
    cs = max(float(st["soundSpeed"][:nInt].max()), 1e-30)
    v = max(float(np.abs(st["velocity"][:nInt]).max()), cs)
    rho = float(st["massDensity"][:nInt].mean())
    P = max(float(np.abs(st["pressure"][:nInt]).max()), 1e-30)
    return dict(DxDt=v, DrhoDt=rho*v/h, DvDt=cs*cs/h, DepsDt=cs*cs*v/h, DvDx=v/h, localDvDx=v/h, gradRho=rho/h, M=1.0,
                localM=1.0, rhoSum=rho, normalization=1.0, maxViscousPressure=P, effViscousPressure=P, XSPHWeightSum=1.0,
                XSPHDeltaV=v, DHDt=v/(h*h), Hideal=1.0/h, massZerothMoment=1.0, massFirstMoment=1.0)


# ---- CheapSynchronousRK2 driven through the oracle (the checker of spheral_b200/integrator.py) -----------------------------
class OracleRK2:
    """Same stage sequence as CheapSynchronousRK2.cc:40-132, every piece an oracle call.  Optional reflecting planes
    (ghosts regenerated every step by nodegen.reflect_ghosts, refreshed by nodegen.reflect_apply)."""

    def __init__(self, orc, oo, so, OT, st, densityUpdate=1, gradhCorrection=True, dtMin=0.0, dtMax=1.0e100, dtGrowth=2.0,
                 planes=None, nInt=None, crk=False, volume_policy=True):
        self.crk = crk
        self.volume_policy = volume_policy
        if crk:
            gradhCorrection = False
        self.orc, self.oo, self.so, self.OT = orc, oo, so, OT
        self.ndim = oo.ndim
        self.N = st["position"].shape[0] if nInt is None else nInt
        self.s = {k: np.array(v[:self.N], dtype=np.float64, copy=True) for k, v in to_oracle_state(st).items()}
        self.s["eps"] = np.array(st["specificThermalEnergy"][:self.N], dtype=np.float64, copy=True)
        self.nGhost, self.planes, self.ctl = 0, planes, None
        self.densityUpdate, self.gradhCorrection = densityUpdate, gradhCorrection
        self.dtMin, self.dtMax, self.dtGrowth = dtMin, dtMax, dtGrowth
        self.t, self.cycle, self.lastDt = 0.0, 0, 1.0e100
        self.needQ = (oo.Qkind == orc.Q_LIMITED_MG) or bool(oo.balsara)
        self.derivs = None
        self.reason, self.node = "", 0

    # -- boundaries ------------------------------------------------------------------------------------------------------
    def _set_ghosts(self):
        if not self.planes:
            return
        inner = {k: v[:self.N] for k, v in self.s.items()}
        out, self.ctl, n0 = ng.reflect_ghosts(self.ndim, inner, self.planes, self.OT.kext, per_plane=True)
        self.s = out
        self.nGhost = out["pos"].shape[0] - self.N

    def _apply_ghosts(self):
        if self.planes:
            ng.reflect_apply(self.ndim, self.s, self.planes, self.ctl, self.N)

    def _enforce(self):
        if not self.planes:
            return
        for point, normal in self.planes:
            nhat = np.asarray(normal, dtype=float)/np.linalg.norm(normal)
            sd = (self.s["pos"][:self.N] - np.asarray(point, dtype=float)) @ nhat
            bad = np.nonzero(sd < 0.0)[0]
            self.s["pos"][bad] -= 2.0*np.outer(sd[bad], nhat)
            self.s["vel"][bad] -= 2.0*np.outer(self.s["vel"][bad] @ nhat, nhat)
            if len(bad):      # PlanarBoundary::updateViolationNodes also enforces the boundary on H (PlanarBoundary.cc:193-195)
                self.s["H"][bad] = ng.reflect_map(self.ndim, "H", self.s["H"][bad], sd[bad], nhat)

    # -- pieces ------------------------------------------------------------------------------------------------------------
    def _pairs(self):
        self.pi, self.pj, self.cnt = self.orc.pairs(self.ndim, self.N, self.nGhost, self.s["pos"], self.s["H"], self.OT.kext)

    def _crk_volume(self):
        v0 = self.s.get("vol")
        if v0 is None or v0.shape[0] != self.N + self.nGhost:
            v0 = self.s["mass"]/self.s["rho"]
        self.s["vol"] = self.orc.crk_sum_volume(self.ndim, self.OT, self.N, self.nGhost, self.s["pos"], self.s["H"], self.pi, self.pj, vol=v0)
        self._apply_ghosts()                      # RKCorrections::preStepInitialize applies the boundaries to the volume (RKCorrections.cc:298-340)

    def _crk_corrections(self):
        if not self.crk:
            return
        c0 = self.s.get("corr")
        if c0 is None or c0.shape[0] != self.N + self.nGhost:
            c0 = np.zeros((self.N + self.nGhost, (self.ndim + 1)**2)); c0[:, 0] = 1.0
        self.s["corr"] = self.orc.crk_corrections(self.ndim, self.OT, self.N, self.nGhost, self.s["pos"], self.s["H"], self.s["vol"],
                                                  self.pi, self.pj, corr=c0)
        self._apply_ghosts()                      # RKCorrections::initialize applies the boundaries to the corrections (RKCorrections.cc:346-372)

    def _sum_density(self):
        if self.crk:
            self.s["rho"] = self.orc.crk_sum_density(self.ndim, self.OT, self.N, self.nGhost, self.s["pos"], self.s["mass"], self.s["vol"],
                                                     self.s["H"], self.pi, self.pj, rhoMin=self.so.rhoMin, rhoMax=self.so.rhoMax,
                                                     rho=self.s["rho"])
            return
        self.s["rho"] = self.orc.sum_mass_density(self.ndim, self.OT, self.N, self.nGhost, self.s["pos"], self.s["mass"], self.s["H"],
                                                  self.pi, self.pj, rho=self.s["rho"])

    def _eos(self):
        self.s["P"], self.s["cs"] = self.orc.eos_gamma_law(self.so, self.s["rho"], self.s["eps"])

    def _omega(self):
        self.s["omega"] = self.orc.omega_gradh(self.ndim, self.OT, self.N, self.nGhost, self.s["pos"], self.s["H"], self.pi, self.pj,
                                               self.cnt, omega=self.s["omega"])

    def _evaluate(self):
        if self.crk:
            self.derivs = self.orc.crk_evaluate_derivatives(self.oo, self.OT, self.s, self.s["vol"], self.s["corr"], self.N, self.nGhost,
                                                            self.pi, self.pj)
        else:
            self.derivs = self.orc.evaluate_derivatives(self.oo, self.OT, self.s, self.N, self.nGhost, self.pi, self.pj, self.cnt)
        self.pairs_eval = (self.pi, self.pj)
        if self.planes and self.oo.compatibleEnergy:
            # SPHBase::finalizeDerivatives (SPHBase.cc:502-519): ghost values of the acceleration and the energy derivative,
            # which SpecificThermalEnergyPolicy reads for the ghost end of an internal-ghost pair
            f = dict(pos=self.s["pos"], DvDt=self.derivs["DvDt"], DepsDt=self.derivs["DepsDt"])
            ng.reflect_apply(self.ndim, f, self.planes, self.ctl, self.N)

    def _post_state_update(self):
        if self.needQ:
            q = np.zeros((self.N + self.nGhost, self.ndim*self.ndim))
            q[:self.N] = self.derivs["DvDx"][:self.N]
            self.s["DvDxQ"] = q
            self._apply_ghosts()
        if self.gradhCorrection:
            self._omega()
            self._apply_ghosts()

    def _update(self, mult, timeAdvanceOnly):
        d = self.derivs
        epsDone = False
        if self.oo.compatibleEnergy and not timeAdvanceOnly:
            pi, pj = self.pairs_eval
            self.s["eps"] = self.orc.update_energy_compatible(self.ndim, self.N, self.nGhost, self.s["mass"], self.s["vel"], d["DvDt"],
                                                              d["DepsDt"], pi, pj, d["pairAccelerations"], mult, self.s["eps"])
            epsDone = True
        out = self.orc.state_update(self.oo, self.so, self.N, self.nGhost, mult, timeAdvanceOnly, d, self.s, epsDone=epsDone)
        self.s.update(out)
        if self.crk and self.volume_policy:
            # ContinuityVolumePolicy (RK/ContinuityVolumePolicy.cc:33-66; CRKSPHBase.cc:155 enrolls the volume with it): depends on mass
            # and mass density, so it fires after the density update; the same in timeAdvanceOnly mode (UpdatePolicyBase.hh:54-61)
            N, m, rho = self.N, self.s["mass"][:self.N], self.s["rho"][:self.N]
            Hdet = np.linalg.det(ng.sym_to_full(self.ndim, self.s["H"][:N]))
            volMax = {1: 2.0, 2: np.pi, 3: 4.0*np.pi/3.0}[self.ndim]/Hdet
            inv = lambda x: np.where(x < 0.0, -1.0, 1.0)/np.maximum(1.0e-30, np.abs(x))     # safeInvVar (Utilities/safeInv.hh:24-27)
            dVdt = -m*inv(rho*rho)*np.asarray(d["DrhoDt"])[:N]
            vol = np.array(self.s["vol"], copy=True)
            vol[:N] = np.maximum(0.5*m*inv(rho), np.minimum(volMax, vol[:N] + mult*dVdt))
            self.s["vol"] = vol

    def _select_dt(self, maxTime):
        vote, why, node = self.orc.hydro_dt(self.oo, self.so, self.N, self.s["vel"], self.s["H"], self.s["rho"], self.s["cs"],
                                            self.derivs, self.pi, self.pj)
        dtMin, dtMax = min(self.dtMin, maxTime - self.t), min(self.dtMax, maxTime - self.t)
        dt = dtMax
        if 0.0 < vote < dt:
            dt, self.reason, self.node = vote, why, node
        dt = min(dt, self.dtGrowth*self.lastDt)
        return min(dtMax, max(dtMin, dt))

    def _pad_derivs(self):
        """Derivatives of the previous step live on the internal nodes; the ghost set may have changed since."""
        n = self.N + self.nGhost
        for k, v in list(self.derivs.items()):
            if k == "pairAccelerations" or v.shape[0] == n:
                continue
            w = np.zeros((n,) + v.shape[1:])
            w[:self.N] = v[:self.N]
            self.derivs[k] = w

    def initializeDerivatives(self):
        self._set_ghosts()
        self._pairs()
        if self.crk:
            self._crk_volume()
        if self.densityUpdate == 1:
            self._sum_density()
        self._eos()
        if self.gradhCorrection:
            self._omega()
        self._apply_ghosts()
        self._crk_corrections()
        self._evaluate()

    def step(self, maxTime=1.0e100):
        self._set_ghosts()
        self._pad_derivs()
        self._pairs()
        if self.crk:
            self._crk_volume()
        if self.densityUpdate == 1:
            self._sum_density()
            self._eos()
            self._apply_ghosts()
        self._crk_corrections()
        dt = self._select_dt(maxTime)
        hdt = 0.5*dt
        s0 = {k: np.array(v, copy=True) for k, v in self.s.items()}
        self._update(hdt, True)
        self._apply_ghosts()
        self._post_state_update()
        self._crk_corrections()
        self._evaluate()
        self.s = s0                 # state.assign(state0) restores every registered field, the Q gradient and omega included
        self._update(dt, False)
        self.t += dt
        self._apply_ghosts()
        self._post_state_update()
        self._enforce()
        self.cycle += 1
        self.lastDt = dt
        return dt

    def total_energy(self):
        m, v, e = self.s["mass"][:self.N], self.s["vel"][:self.N], self.s["eps"][:self.N]
        return float(np.sum(m*(0.5*np.sum(v*v, axis=1) + e)))
