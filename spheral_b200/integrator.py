"""CheapSynchronousRK2 -- host-side mirror of Spheral's integrator over the device-resident state.

Reference: Integrator/CheapSynchronousRK2.cc:40-132 (the stage sequence), Integrator/Integrator.cc:66-111 (step: connectivity
update, retry with a halved dt multiplier), :114-166 (selectDt: package votes, dtGrowth*lastDt, [dtMin, dtMax]), and the hooks
of the SPH package that run inside a stage: SPHBase::preStepInitialize (sum density, SPHBase.cc:322-352),
SPHBase::postStateUpdate (grad-h correction, :527-547), ArtificialViscosityHandle::postStateUpdate (DvDx -> Q gradient).

Every field stays on the GPU: one step issues the C-ABI calls below and reads back a single number, the hydro time-step vote
(sphb200_compute_dt).  Method and attribute names follow the reference (`step`, `advance`, `currentTime`, `currentCycle`,
`lastDt`, `dtMin`, `dtMax`, `dtGrowth`, `allowDtCheck`).
"""
from . import _lib as L
from . import engine as E

INTEGRATE_DENSITY, RIGOROUS_SUM_DENSITY = 0, 1      # MassDensityType (Hydro/GenericHydro.hh)


class CheapSynchronousRK2:
    def __init__(self, engine, step_options=None, densityUpdate=RIGOROUS_SUM_DENSITY, gradhCorrection=True, cfl=0.25,
                 useVelocityMagnitudeForDt=False, dtMin=0.0, dtMax=1.0e100, dtGrowth=2.0, allowDtCheck=False,
                 ghostRefresh=None, reflectingPlanes=None, distributed=None, boundaries=None, lazyOmega=False):
        self.engine = engine
        # lazyOmega: the grad-h correction recomputed at the END of a step (SPHBase::postStateUpdate after the full update,
        # CheapSynchronousRK2.cc:115) is never read by an evaluation -- the next step recomputes it on its trial state before
        # evaluateDerivatives and the dt vote does not use it.  With lazyOmega it is computed only when somebody looks
        # (ensureOmega(), dumpState()), on the connectivity of the step that produced the state, as the reference does.  Off by
        # default: engine.download_state("omegaGradh") between steps then needs ensureOmega() first.
        self.lazyOmega, self._omegaStale = bool(lazyOmega), False
        self.so = step_options if step_options is not None else E.make_step_options()
        self.densityUpdate, self.gradhCorrection = densityUpdate, gradhCorrection
        self.cfl, self.useVelocityMagnitudeForDt = cfl, useVelocityMagnitudeForDt
        self.dtMin, self.dtMax, self.dtGrowth = dtMin, dtMax, dtGrowth
        self.allowDtCheck = allowDtCheck
        # applyGhostBoundaries + finalizeGhostBoundaries: a callable refreshing the ghost entries on the device (halo exchange);
        # None for a problem without ghost nodes
        self.ghostRefresh = ghostRefresh
        # reflecting planes handled on the device (ReflectingBoundary): list of (point, inward normal)
        # `boundaries` is the general form (engine.boundary_configure: reflecting and periodic planar boundaries)
        self.reflectingPlanes = reflectingPlanes or boundaries
        if boundaries:
            if reflectingPlanes:
                raise ValueError("give the planes either as reflectingPlanes or inside boundaries, not both")
            engine.boundary_configure(boundaries)
        elif reflectingPlanes:
            engine.reflect_configure(reflectingPlanes)
        # domain decomposition (spheral_b200.distributed.DistributedSPH): ghost exchange with the neighbouring slabs over NCCL
        # With planes AND slabs the ghost tail is [plane ghosts | halo]: the planes come first in the boundary list, the distributed
        # boundary last, and it exchanges the plane ghosts near a slab face as well (Integrator.cc:389-445, SpheralController.py:
        # the DistributedBoundary is appended after the problem's own boundaries).
        self.distributed = distributed
        self.currentTime, self.currentCycle, self.lastDt = 0.0, 0, 1.0e100
        self.dtMultiplier = 1.0
        self.lastDtReason, self.lastDtNode = "", 0
        self._needQ = (engine.options.Qkind == L.Q_LIMITED_MG) or bool(engine.options.balsara)
        # CRKSPH (engine created with hydro = HYDRO_CRKSPH): the RKCorrections package runs in front of the hydro
        # (SpheralController.py:690-745): volumes in preStepInitialize, corrections in initialize (before EVERY evaluation)
        self._crk = engine.options.hydro == L.HYDRO_CRKSPH
        if self._crk:
            self.gradhCorrection = False                     # CRKSPHBase has no grad-h term

    # -- pieces of a stage -----------------------------------------------------------------------------------------------
    def _ghosts(self):
        """applyGhostBoundaries + finalizeGhostBoundaries (Integrator.cc:530-600)."""
        if self.reflectingPlanes:
            self.engine.reflect_apply_ghosts()
        if self.distributed is not None:
            self.distributed.apply_ghosts()
        if self.ghostRefresh is not None:
            self.ghostRefresh()

    def _finalize_derivatives(self):
        """SPHBase::finalizeDerivatives (SPHBase.cc:502-519): boundary conditions on DvDt and DepsDt for the compatible energy."""
        if not self.engine.options.compatibleEnergy:
            return
        if self.reflectingPlanes:
            self.engine.reflect_finalize_derivatives()
        if self.distributed is not None:
            self.distributed.finalize_derivatives()

    def _set_ghost_nodes(self):
        """Integrator::setGhostNodes (Integrator.cc:372-445): regenerate the ghost nodes, then refresh their values."""
        nPlaneGhosts = self.engine.reflect_set_ghost_nodes() if self.reflectingPlanes else 0
        if self.distributed is not None:
            self.distributed.refresh_ghosts(build=False, boundary_ghosts=nPlaneGhosts)
        if self.ghostRefresh is not None:
            self.ghostRefresh()

    def _post_state_update(self, endOfStep=False):
        """Integrator::postStateUpdate (Integrator.cc:252-271)."""
        e = self.engine
        if self._needQ:
            e.copy_DvDx_to_Q()                       # ArtificialViscosityHandle.cc:165-180
        if self.gradhCorrection:
            if endOfStep and self.lazyOmega:
                self._omegaStale = True
                return
            e.compute_omega_gradh()                  # SPHBase.cc:539-547
            self._omegaStale = False
            self._ghosts()

    def ensureOmega(self):
        """lazyOmega: bring the grad-h correction of the current state up to date (no-op otherwise).  Must be called before the
        next step rebuilds the connectivity to reproduce the reference's end-of-step value."""
        if self._omegaStale:
            self.engine.compute_omega_gradh()
            self._omegaStale = False
            self._ghosts()

    def _pre_step_initialize(self):
        """Integrator::preStepInitialize (Integrator.cc:177-183): RKCorrections::preStepInitialize (RK/RKCorrections.cc:298-340)
        then the hydro's (SPHBase.cc:322-352 / CRKSPHBase.cc:229-256); returns True if the density was replaced."""
        e, so = self.engine, self.so
        if self._crk:
            e.crk_compute_volume()
            if self.distributed is not None:
                self.distributed.mark_ready("volume")
            self._ghosts()
        if self.densityUpdate == RIGOROUS_SUM_DENSITY:
            if self._crk:
                e.crk_sum_mass_density(so.rhoMin, so.rhoMax)
            else:
                e.sum_mass_density()
            e.update_eos_gamma_law(so.eos)               # pressure / sound speed follow the density they depend on
            self._ghosts()
            return True
        return False

    def _initialize(self):
        """Integrator::initializeDerivatives (Integrator.cc:186-210): RKCorrections::initialize (RKCorrections.cc:346-372)."""
        if self._crk:
            self.engine.crk_compute_corrections()
            if self.distributed is not None:
                self.distributed.mark_ready("rkCorrections")
            self._ghosts()

    def selectDt(self, dtMin, dtMax):
        """Integrator::selectDt (Integrator.cc:114-166) with the one package vote of GenericHydro::dt."""
        vote, why, node = self.engine.compute_dt(self.cfl, self.useVelocityMagnitudeForDt)
        if self.distributed is not None:
            vote = self.distributed.allreduce_min(vote)          # allReduce(dt, MIN), Integrator.cc:150-154
        dt = dtMax
        if 0.0 < vote < dt:
            dt = vote
            self.lastDtReason, self.lastDtNode = why, node
        dt *= self.dtMultiplier
        dt = min(dt, self.dtGrowth*self.lastDt)
        return min(dtMax, max(dtMin, dt))

    def initializeDerivatives(self):
        """The first evaluation of a run: CheapSynchronousRK2 advances the trial state with the previous step's derivatives,
        so a fresh problem needs one (the reference does this in SpheralController.reinitializeProblem -> evaluateDerivatives)."""
        e = self.engine
        self._set_ghost_nodes()
        e.build_pairs()
        if not self._pre_step_initialize():
            e.update_eos_gamma_law(self.so.eos)
        if self.gradhCorrection:
            e.compute_omega_gradh()
        self._ghosts()
        self._initialize()
        e.evaluate_derivatives(self.currentTime, 0.0)
        self._finalize_derivatives()

    # -- one step ----------------------------------------------------------------------------------------------------------
    def _try_step(self, maxTime):
        """CheapSynchronousRK2::step(maxTime, state, derivs) (CheapSynchronousRK2.cc:40-132)."""
        e, so = self.engine, self.so
        t = self.currentTime
        self._pre_step_initialize()
        self._initialize()                            # initializeDerivatives(t, 0) (CheapSynchronousRK2.cc:56)
        dt = self.selectDt(min(self.dtMin, maxTime - t), min(self.dtMax, maxTime - t))
        hdt = 0.5*dt
        e.state_copy()                                # state0
        # trial advance to the mid point with the previous derivatives (timeAdvanceOnly)
        e.state_update(so, hdt, True)
        self._ghosts()
        self._post_state_update()
        # derivatives at the mid point, on the connectivity of the step start
        self._initialize()                            # initializeDerivatives(t + hdt, hdt) (CheapSynchronousRK2.cc:86)
        e.evaluate_derivatives(t + hdt, hdt)
        self._finalize_derivatives()
        if self.allowDtCheck:
            dtnew = self.selectDt(min(self.dtMin, maxTime - t), min(self.dtMax, maxTime - t))
            if dtnew < 0.5*dt:
                e.state_assign()
                return False
        # full step from state0 with the mid-point derivatives
        e.state_assign()
        e.state_update(so, dt, False)
        self.currentTime = t + dt
        self._ghosts()
        self._post_state_update(endOfStep=True)
        if self.reflectingPlanes:
            e.reflect_enforce()                       # enforceBoundaries (CheapSynchronousRK2.cc:121-123)
        self.currentCycle += 1
        self.lastDt = dt
        return True

    def step(self, maxTime=1.0e100):
        """Integrator::step(maxTime) (Integrator.cc:66-111): neighbour update, then up to 10 attempts with a halved dt."""
        self._omegaStale = False                      # a stale end-of-step grad-h correction dies here, unread (lazyOmega)
        self._set_ghost_nodes()                       # setGhostNodes
        self.engine.build_pairs()                     # Neighbor::updateNodes + ConnectivityMap::computeConnectivity
        ok, count = False, 0
        allow = self.allowDtCheck
        while not ok and count < 10:
            count += 1
            if count == 10:
                self.allowDtCheck = False
            ok = self._try_step(maxTime)
            if not ok:
                self.dtMultiplier *= 0.5
        self.allowDtCheck = allow
        self.dtMultiplier = 1.0
        return ok

    # -- restart (RestartableObject: Integrator::dumpState / restoreState, SPHBase::dumpState / restoreState) -------------------------
    RESTART_DERIVS = ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "maxViscousPressure", "DHDt", "Hideal")

    def dumpState(self):
        """Everything a restart needs, as host arrays: the state fields on the device, the derivative fields the next trial
        advance and dt vote read (SPHBase.cc:714-735 dumps them for the same reason), and the integrator's counters
        (Integrator.cc dumpState: time, cycle, lastDt)."""
        e = self.engine
        self.ensureOmega()
        names = [k for k in L.STATE_FIELDS if k not in ("fCl", "fCq", "volume", "rkCorrections") or (self._crk and k in ("volume", "rkCorrections"))]
        have = []
        for k in names:
            try:
                e.download_state(k)
                have.append(k)
            except E.SPHB200Error:
                pass
        return dict(state=e.download_state(*have), derivs=e.download_derivs(*self.RESTART_DERIVS) if self._derivs_downloadable() else None,
                    nInternal=e.nInternal, nGhost=e.nGhost, time=self.currentTime, cycle=self.currentCycle, lastDt=self.lastDt)

    def _derivs_downloadable(self):
        try:
            self.engine.download_derivs("DrhoDt")
            return True
        except E.SPHB200Error:
            return False

    def restoreState(self, dump):
        e = self.engine
        e.set_nodes(dump["nInternal"], dump["nGhost"])
        e.upload_state(**dump["state"])
        if dump["derivs"] is not None:
            e.upload_derivs(**dump["derivs"])
        self.currentTime, self.currentCycle, self.lastDt = dump["time"], dump["cycle"], dump["lastDt"]

    def advance(self, goalTime, maxSteps=None):
        n = 0
        while self.currentTime < goalTime and (maxSteps is None or n < maxSteps):
            self.step(goalTime)
            n += 1
        return n


def iterateIdealH(engine, maxIterations=100, tolerance=1.0e-10, setGhostNodes=None):
    """Utilities/iterateIdealH.cc: the start-up relaxation of the smoothing scales, with the state on the device.  Every
    iteration regenerates the ghost nodes (setGhostNodes: e.g. engine.reflect_set_ghost_nodes), rebuilds the connectivity,
    evaluates the smoothing-scale derivatives and replaces H by the ideal H on the nodes that have not converged.
    Returns (iterations, maxDeltaH).  SPH and classic ASPH (hEvolution = H_ASPH_CLASSIC) smoothing scales."""
    it, maxDeltaH = 0, 2.0*tolerance
    while it < maxIterations and maxDeltaH > tolerance:
        it += 1
        if setGhostNodes is not None:
            setGhostNodes()
        engine.build_pairs()
        engine.evaluate_derivatives(0.0, 1.0)
        maxDeltaH = engine.iterate_ideal_h(it == 1, tolerance)
    if setGhostNodes is not None:
        setGhostNodes()
    return it, maxDeltaH
