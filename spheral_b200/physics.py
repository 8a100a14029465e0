"""Host-side mirror of the reference's Physics package interface for the SPH hot path.

A Spheral script does

    db = DataBase(); db.appendNodeList(nodes)
    hydro = SPH(dataBase=db, W=WT, ...)                       # SPH/SPHHydros.py:9-140
    hydro.initializeProblemStartup(db)
    state = State(db, [hydro]); derivs = StateDerivatives(db, [hydro])
    hydro.evaluateDerivatives(t, dt, db, state, derivs)        # Physics/Physics.hh:44-52

and this module provides exactly those names with the same argument meaning; the work is done by libsphb200 (CUDA)
through the C ABI of include/sphb200.h.  Only what the hot path touches is mirrored:

    reference                                               here
    ------------------------------------------------------  -----------------------------------------------------
    NodeList / FluidNodeList (NodeList/FluidNodeList.hh)     FluidNodeList: named AoS numpy fields, internal+ghost
    DataBase (DataBase/DataBase.hh)                          DataBase: list of NodeLists (one fluid NodeList)
    State / StateDerivatives (DataBase/StateBase.hh:158-170) dict keyed "<field name>|<NodeList name>"
    Physics<Dim> (Physics/Physics.hh:26-235)                 Physics base with the same hook names
    SPH<Dim> (SPH/SPH.hh:40-56, SPH.cc:85-555)               SPHB200 (+ the SPH()/ASPH() factories)
    ArtificialViscosityHandle (ArtificialViscosity/*.hh)     MonaghanGingoldViscosity / LimitedMonaghanGingoldViscosity
    SPHSmoothingScale / ASPHSmoothingScale                   same names (configuration carriers, post sub-packages)
    PairwiseField "pair-wise accelerations" (SPH.cc:129-133) lazy device view, hydro.pairAccelerations

There is no CPU implementation behind any of this: every compute call goes to the GPU library or raises.
"""
import numpy as np

from . import _lib as L
from .engine import Engine, SPHB200Error, make_options


# ---- HydroFieldNames (Hydro/HydroFieldNames.cc:9-66) ----------------------------------------------------------------------
class HydroFieldNames:
    mass = "mass"
    position = "position"
    velocity = "velocity"
    H = "H"
    massDensity = "mass density"
    specificThermalEnergy = "specific thermal energy"
    pressure = "pressure"
    soundSpeed = "sound speed"
    omegaGradh = "grad h corrections"
    timeStepMask = "time step mask"
    velocityGradient = "velocity gradient"
    internalVelocityGradient = "internal velocity gradient"
    hydroAcceleration = "delta velocity hydro"
    normalization = "normalization"
    maxViscousPressure = "max viscous pressure"
    effectiveViscousPressure = "effective viscous pressure"
    massDensityCorrection = "density summation correction"
    XSPHDeltaV = "XSPH delta vi"
    XSPHWeightSum = "XSPH weight sum"
    massZerothMoment = "mass zeroth moment"
    massFirstMoment = "mass first moment"
    pairAccelerations = "pair-wise accelerations"
    M_SPHCorrection = "M SPH gradient correction"
    massDensityGradient = "mass density gradient"
    ArtificialViscousClMultiplier = "Cl multiplier for artificial viscosity"
    ArtificialViscousCqMultiplier = "Cq multiplier for artificial viscosity"
    ArtificialViscosityVelocityGradient = "velocity gradient for artificial viscosity"
    volume = "volume"
    surfacePoint = "surface point"


class RKOrder:
    """RK/RKCorrectionParams.hh: the orders this path implements."""
    ZerothOrder, LinearOrder = 0, 1


class RKFieldNames:
    """RK/RKFieldNames.hh: rkCorrections(order) / reproducingKernel(order)."""
    @staticmethod
    def rkCorrections(order): return "rkCorrections_%d" % order
    @staticmethod
    def reproducingKernel(order): return "reproducingKernel_%d" % order


RKSumVolume, RKMassOverDensity, RKVoronoiVolume = "RKSumVolume", "RKMassOverDensity", "RKVoronoiVolume"


DELTA, NEW = "delta ", "new "          # IncrementState::prefix(), ReplaceState::prefix()

# C-ABI name -> State key (inputs) ; SPH.cc:206-216
STATE_KEYS = {
    "position": HydroFieldNames.position, "velocity": HydroFieldNames.velocity, "H": HydroFieldNames.H,
    "mass": HydroFieldNames.mass, "massDensity": HydroFieldNames.massDensity,
    "specificThermalEnergy": HydroFieldNames.specificThermalEnergy, "pressure": HydroFieldNames.pressure,
    "soundSpeed": HydroFieldNames.soundSpeed, "omegaGradh": HydroFieldNames.omegaGradh,
    "DvDxQ": HydroFieldNames.ArtificialViscosityVelocityGradient,
    "fCl": HydroFieldNames.ArtificialViscousClMultiplier, "fCq": HydroFieldNames.ArtificialViscousCqMultiplier,
    "volume": HydroFieldNames.volume, "rkCorrections": RKFieldNames.rkCorrections(RKOrder.LinearOrder),
}
# C-ABI name -> StateDerivatives key (outputs) ; SPH.cc:230-245, SPHBase.cc:127-140, SmoothingScaleBase.cc:45-46
DERIV_KEYS = {
    "DxDt": DELTA + HydroFieldNames.position, "DrhoDt": DELTA + HydroFieldNames.massDensity,
    "DvDt": HydroFieldNames.hydroAcceleration, "DepsDt": DELTA + HydroFieldNames.specificThermalEnergy,
    "DvDx": HydroFieldNames.velocityGradient, "localDvDx": HydroFieldNames.internalVelocityGradient,
    "gradRho": HydroFieldNames.massDensityGradient, "M": HydroFieldNames.M_SPHCorrection,
    "localM": "local " + HydroFieldNames.M_SPHCorrection, "rhoSum": NEW + HydroFieldNames.massDensity,
    "normalization": HydroFieldNames.normalization, "maxViscousPressure": HydroFieldNames.maxViscousPressure,
    "effViscousPressure": HydroFieldNames.effectiveViscousPressure, "XSPHWeightSum": HydroFieldNames.XSPHWeightSum,
    "XSPHDeltaV": HydroFieldNames.XSPHDeltaV, "DHDt": DELTA + HydroFieldNames.H, "Hideal": NEW + HydroFieldNames.H,
    "massZerothMoment": HydroFieldNames.massZerothMoment, "massFirstMoment": HydroFieldNames.massFirstMoment,
}


def _key(field, nodeListName):
    """StateBase::buildFieldKey (DataBase/StateBaseInline.hh:220-227)."""
    return field + "|" + nodeListName


# ---- data model stand-ins ---------------------------------------------------------------------------------------------------
class FluidNodeList:
    """Fields are AoS float64 arrays of length numNodes = numInternalNodes + numGhostNodes (Field/Field.hh:213)."""

    def __init__(self, name, ndim, numInternal=0, numGhost=0, nPerh=2.01, hmin=1.0e-20, hmax=1.0e20,
                 rhoMin=1.0e-10, rhoMax=1.0e10):
        self.name, self.ndim = name, ndim
        self.nodesPerSmoothingScale, self.hmin, self.hmax, self.rhoMin, self.rhoMax = nPerh, hmin, hmax, rhoMin, rhoMax
        self.numInternalNodes, self.numGhostNodes = numInternal, numGhost
        self._fields = {}
        self._version = 0                 # bumped whenever a field array is replaced
        for abi in ("position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy"):
            self._fields[STATE_KEYS[abi]] = np.zeros(self._shape(L.state_width(ndim, abi)))

    @property
    def numNodes(self):
        return self.numInternalNodes + self.numGhostNodes

    def _shape(self, w):
        return (self.numNodes, w) if w > 1 else (self.numNodes,)

    def resize(self, numInternal, numGhost=0):
        self.numInternalNodes, self.numGhostNodes = numInternal, numGhost
        for k, v in list(self._fields.items()):
            w = v.shape[1] if v.ndim > 1 else 1
            self._fields[k] = np.zeros(self._shape(w))
        self._version += 1

    def field(self, name):
        return self._fields[name]

    def setField(self, name, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        if a.shape[0] != self.numNodes:
            raise ValueError("field '%s': %d values for %d nodes" % (name, a.shape[0], self.numNodes))
        self._fields[name] = a
        self._version += 1

    # the accessors a Spheral script uses (NodeList.hh / FluidNodeList.hh)
    def positions(self): return self._fields[HydroFieldNames.position]
    def velocity(self): return self._fields[HydroFieldNames.velocity]
    def Hfield(self): return self._fields[HydroFieldNames.H]
    def mass(self): return self._fields[HydroFieldNames.mass]
    def massDensity(self): return self._fields[HydroFieldNames.massDensity]
    def specificThermalEnergy(self): return self._fields[HydroFieldNames.specificThermalEnergy]


class DataBase:
    """DataBase<Dim>: the list of NodeLists (this path supports one fluid NodeList, as configs C1-C3 use)."""

    def __init__(self):
        self.nodeLists = []
        self.connectivityValid = False

    def appendNodeList(self, nl):
        if self.nodeLists:
            raise SPHB200Error("SPHB200 supports a single fluid NodeList per DataBase")
        self.nodeLists.append(nl)

    @property
    def nDim(self): return self.nodeLists[0].ndim
    @property
    def numFluidNodeLists(self): return len(self.nodeLists)
    @property
    def numSolidNodeLists(self): return 0
    @property
    def numInternalNodes(self): return sum(n.numInternalNodes for n in self.nodeLists)


class StateBase(dict):
    """map "<field>|<NodeList>" -> array (DataBase/StateBase.hh:158-170).  Arrays are REFERENCES to the owners' storage."""

    def enroll(self, nodeListName, field, array):
        self[_key(field, nodeListName)] = array

    def registered(self, field, nodeListName=None):
        if nodeListName is not None:
            return _key(field, nodeListName) in self
        return any(k.split("|")[0] == field for k in self)

    def field(self, field, nodeListName):
        return self[_key(field, nodeListName)]

    def fields(self, field):
        """FieldList of every NodeList's field of that name (StateBase::fields)."""
        return [v for k, v in self.items() if k.split("|")[0] == field]


class State(StateBase):
    def __init__(self, dataBase=None, packages=()):
        super().__init__()
        self.timeAdvanceOnly = False
        for p in packages:
            p.registerState(dataBase, self)


class StateDerivatives(StateBase):
    def __init__(self, dataBase=None, packages=()):
        super().__init__()
        for p in packages:
            p.registerDerivatives(dataBase, self)

    def Zero(self):
        """StateDerivatives::Zero (CheapSynchronousRK2.cc:87)."""
        for v in self.values():
            if isinstance(v, np.ndarray):
                v[...] = 0.0


# ---- package configuration carriers ------------------------------------------------------------------------------------------
class ArtificialViscosityHandle:
    """ArtificialViscosityHandle(Clinear, Cquadratic, kernel) -- ArtificialViscosity/ArtificialViscosityHandle.cc:37-53."""
    Qkind = L.Q_MG

    def __init__(self, Clinear=1.0, Cquadratic=1.0, kernel=None, linearInExpansion=False, quadraticInExpansion=False,
                 **extra):
        self.Cl, self.Cq, self.kernel = Clinear, Cquadratic, kernel
        self.linearInExpansion, self.quadraticInExpansion = linearInExpansion, quadraticInExpansion
        self.epsilon2, self.negligibleSoundSpeed, self.balsaraShearCorrection = 1.0e-2, 1.0e-10, False
        self.etaCritFrac, self.etaFoldFrac = extra.get("etaCritFrac", 1.0), extra.get("etaFoldFrac", 0.2)

    def label(self): return type(self).__name__


class MonaghanGingoldViscosity(ArtificialViscosityHandle):
    """MonaghanGingoldViscosity.cc:41-101"""
    Qkind = L.Q_MG


class LimitedMonaghanGingoldViscosity(ArtificialViscosityHandle):
    """LimitedMonaghanGingoldViscosity.cc:120-219 -- needs the velocity gradient of the previous evaluation."""
    Qkind = L.Q_LIMITED_MG


class SPHSmoothingScale:
    """SPHSmoothingScale(HUpdate, W) -- SmoothingScale/SPHSmoothingScale.cc"""
    hEvolution = L.H_SPH

    def __init__(self, HUpdate=None, W=None):
        self.HEvolution, self.WT = HUpdate, W

    def label(self): return type(self).__name__


class ASPHSmoothingScale(SPHSmoothingScale):
    """ASPHSmoothingScale(HUpdate, W) -- SmoothingScale/ASPHSmoothingScale.cc"""
    hEvolution = L.H_ASPH


class ASPHClassicSmoothingScale(SPHSmoothingScale):
    """ASPHClassicSmoothingScale(HUpdate, W) -- SmoothingScale/ASPHClassicSmoothingScale.cc: the ASPH tensor derivative plus the
    second-moment ideal H (the factories select it with ASPH = "Classic", SPHHydros.py:133-134)"""
    hEvolution = L.H_ASPH_CLASSIC


IdealH, IntegrateH, FixedH = "IdealH", "IntegrateH", "FixedH"
RigorousSumDensity, IntegrateDensity = "RigorousSumDensity", "IntegrateDensity"


class PairAccelerationsView:
    """Lazy device view of PairwiseField<Dim,Vector> "pair-wise accelerations" (Neighbor/PairwiseField.hh:27-80): the values
    stay in HBM until a host reader asks for them; they are invalid after a topology change, like the reference's weak_ptr."""

    def __init__(self, hydro):
        self._hydro = hydro

    def __len__(self):
        return self._hydro._engine.npairs

    def __array__(self, dtype=None, copy=None):
        a = self._hydro._engine.download_pair_accelerations()
        return a if dtype is None else a.astype(dtype)


# ---- Physics<Dim> (Physics/Physics.hh:26-235) ---------------------------------------------------------------------------------
class Physics:
    def __init__(self):
        self._pre, self._post, self._boundaries = [], [], []

    def prependSubPackage(self, p): self._pre.append(p)
    def appendSubPackage(self, p): self._post.append(p)
    def preSubPackages(self): return list(self._pre)
    def postSubPackages(self): return list(self._post)
    def appendBoundary(self, b): self._boundaries.append(b)
    def boundaryConditions(self): return list(self._boundaries)
    def requireConnectivity(self): return True
    def requireGhostConnectivity(self): return False
    def requireOverlapConnectivity(self): return False
    def requireIntersectionConnectivity(self): return False
    def requireVoronoiCells(self): return False
    def requireReproducingKernels(self): return False
    def initializeProblemStartup(self, dataBase): pass
    def initializeProblemStartupDependencies(self, dataBase, state, derivs): pass
    def preStepInitialize(self, dataBase, state, derivs): pass
    def initialize(self, time, dt, dataBase, state, derivs): return False
    def finalizeDerivatives(self, time, dt, dataBase, state, derivs): pass
    def postStateUpdate(self, time, dt, dataBase, state, derivs): return False
    def finalize(self, time, dt, dataBase, state, derivs): pass
    def applyGhostBoundaries(self, state, derivs): pass
    def enforceBoundaries(self, state, derivs): pass
    def extraEnergy(self): return 0.0
    def extraMomentum(self): return None


class SPHB200(Physics):
    """Drop-in for SPH<Dim> (SPH/SPH.hh:40-56): same constructor keywords, same hooks, same field keys."""

    def __init__(self, dataBase, Q, W, WPi=None, cfl=0.25, useVelocityMagnitudeForDt=False,
                 compatibleEnergyEvolution=True, evolveTotalEnergy=False, gradhCorrection=True, XSPH=True,
                 correctVelocityGradient=True, sumMassDensityOverAllNodeLists=True, densityUpdate=RigorousSumDensity,
                 epsTensile=0.0, nTensile=4.0, xmin=None, xmax=None, device=0):
        super().__init__()
        if compatibleEnergyEvolution and evolveTotalEnergy:          # SPH.cc:97-98 (VERIFY2)
            raise SPHB200Error("SPH error : you cannot simultaneously use both compatibleEnergyEvolution and evolveTotalEnergy")
        self.Q = Q
        self._W, self._WPi = W, (WPi if WPi is not None else W)
        self.cfl, self.useVelocityMagnitudeForDt = cfl, useVelocityMagnitudeForDt
        self._compat, self._evolveE, self.gradhCorrection = compatibleEnergyEvolution, evolveTotalEnergy, gradhCorrection
        self._XSPH, self._corrVG = XSPH, correctVelocityGradient
        self.sumMassDensityOverAllNodeLists, self.densityUpdate = sumMassDensityOverAllNodeLists, densityUpdate
        self._epsTensile, self._nTensile = epsTensile, nTensile
        self.xmin, self.xmax = xmin, xmax
        self._smoothingScaleMethod = None
        self._device = device
        self._engine = None
        self._own = {}                     # package-owned fields (SPHBase.hh:223-251), name -> array
        self._pairAccelerations = PairAccelerationsView(self)
        nl = dataBase.nodeLists[0]
        self._ndim = nl.ndim

    # -- properties (SPHBase.hh:131-197) ----------------------------------------------------------------------------------
    _hydro = L.HYDRO_SPH

    def label(self): return "SPH"
    kernel = property(lambda s: s._W)
    PiKernel = property(lambda s: s._WPi)
    pairAccelerations = property(lambda s: s._pairAccelerations)

    def _flag(name):                                          # noqa: N805 -- property factory
        def get(self): return getattr(self, name)
        def set_(self, v):
            setattr(self, name, v)
            self._push_options()
        return property(get, set_)
    compatibleEnergyEvolution = _flag("_compat")
    evolveTotalEnergy = _flag("_evolveE")
    XSPH = _flag("_XSPH")
    correctVelocityGradient = _flag("_corrVG")
    epsilonTensile = _flag("_epsTensile")
    nTensile = _flag("_nTensile")
    del _flag

    def _options(self, nl):
        Q, sm = self.Q, self._smoothingScaleMethod
        return make_options(self._ndim, compatibleEnergy=int(self._compat), evolveTotalEnergy=int(self._evolveE),
                            XSPH=int(self._XSPH), correctVelocityGradient=int(self._corrVG), epsTensile=self._epsTensile,
                            nTensile=self._nTensile, nPerh=nl.nodesPerSmoothingScale, Qkind=Q.Qkind, Cl=Q.Cl, Cq=Q.Cq,
                            eps2=Q.epsilon2, negligibleSoundSpeed=Q.negligibleSoundSpeed,
                            balsara=int(Q.balsaraShearCorrection), linearInExpansion=int(Q.linearInExpansion),
                            quadraticInExpansion=int(Q.quadraticInExpansion), etaCritFrac=Q.etaCritFrac,
                            etaFoldFrac=Q.etaFoldFrac, hEvolution=(sm.hEvolution if sm is not None else L.H_NONE),
                            hmin=nl.hmin, hmax=nl.hmax, hydro=self._hydro, hminratio=getattr(nl, "hminratio", 0.1))

    def _push_options(self):
        if self._engine is not None:
            o = self._options(self._nl)
            self._engine.options = o
            self._engine._check(self._engine._lib.sphb200_set_options(self._engine._h, o))

    # -- Physics hooks ---------------------------------------------------------------------------------------------------------
    def initializeProblemStartup(self, dataBase):
        """SPHBase::initializeProblemStartup (SPHBase.cc:116-142): size the package-owned fields; here also create the
        device context and upload the kernel tables."""
        nl = self._nl = dataBase.nodeLists[0]
        self._engine = Engine(self._ndim, device=self._device, options=self._options(nl))
        self._engine.set_kernel_table(self._W, L.TABLE_W)
        if self._WPi is not self._W and not (self._WPi == self._W):
            self._engine.set_kernel_table(self._WPi, L.TABLE_WPI)
        self._resize_owned(nl)

    def _resize_owned(self, nl):
        n, nd = nl.numNodes, self._ndim
        def mk(w): return np.zeros((n, w) if w > 1 else n)
        own = self._own
        for name in (HydroFieldNames.pressure, HydroFieldNames.soundSpeed):
            if name not in own or own[name].shape[0] != n:
                own[name] = mk(1)
        if HydroFieldNames.omegaGradh not in own or own[HydroFieldNames.omegaGradh].shape[0] != n:
            own[HydroFieldNames.omegaGradh] = np.ones(n)                      # SPHBase.cc:123 (omega initialised to 1)
        if HydroFieldNames.timeStepMask not in own or own[HydroFieldNames.timeStepMask].shape[0] != n:
            own[HydroFieldNames.timeStepMask] = np.ones(n, dtype=np.int32)
        if self.Q.Qkind == L.Q_LIMITED_MG or self.Q.balsaraShearCorrection:
            k = HydroFieldNames.ArtificialViscosityVelocityGradient
            if k not in own or own[k].shape[0] != n:
                own[k] = mk(nd*nd)
        for abi, key in DERIV_KEYS.items():
            if key not in own or own[key].shape[0] != n:
                own[key] = mk(L.deriv_width(nd, abi))

    def registerState(self, dataBase, state):
        """SPHBase::registerState + SPH::registerState (SPHBase.cc:203-275, SPH.cc:85-113): NodeList-owned m, r, v, rho, eps, H
        and package-owned P, cs, omega, mask are enrolled by reference."""
        nl = dataBase.nodeLists[0]
        self._resize_owned(nl)
        for abi in ("mass", "position", "velocity", "massDensity", "specificThermalEnergy", "H"):
            state.enroll(nl.name, STATE_KEYS[abi], nl.field(STATE_KEYS[abi]))
        for name in (HydroFieldNames.pressure, HydroFieldNames.soundSpeed, HydroFieldNames.omegaGradh, HydroFieldNames.timeStepMask):
            state.enroll(nl.name, name, self._own[name])
        k = HydroFieldNames.ArtificialViscosityVelocityGradient
        if k in self._own:
            state.enroll(nl.name, k, self._own[k])
        state.policies = getattr(state, "policies", {})
        state.policies[HydroFieldNames.specificThermalEnergy] = (
            "SpecificThermalEnergyPolicy" if self._compat else
            "SpecificFromTotalThermalEnergyPolicy" if self._evolveE else "IncrementState")       # SPH.cc:101-113

    def registerDerivatives(self, dataBase, derivs):
        """SPHBase::registerDerivatives + SPH::registerDerivatives (SPHBase.cc:277-316, SPH.cc:118-135)."""
        nl = dataBase.nodeLists[0]
        self._resize_owned(nl)
        for abi, key in DERIV_KEYS.items():
            derivs.enroll(nl.name, key, self._own[key])
        if self._compat:
            derivs[HydroFieldNames.pairAccelerations] = self._pairAccelerations      # misc key, SPH.cc:132

    def updateConnectivity(self, dataBase, state):
        """What Integrator::setGhostNodes asks for because requireConnectivity() is true (Integrator.cc:372-445):
        Neighbor::updateNodes + DataBase::updateConnectivityMap, on the device."""
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("position", "H"), keep_connectivity=False)
        self._engine.build_pairs()
        dataBase.connectivityValid = True
        return self._engine.npairs

    def _sync_state(self, nl, state, names, keep_connectivity=True):
        """Copy the named state fields from the host State to the device -- always, whatever happened to the host arrays since the
        last call: the State hands out references, so an in-place update by another package is indistinguishable from no update.
        keep_connectivity: the device keeps its pair lists although positions / H are refreshed, which is what the reference does
        between the neighbour update of a step and its mid-step evaluateDerivatives (CheapSynchronousRK2.cc:76-99); only
        updateConnectivity invalidates them."""
        e = self._engine
        if (nl.numInternalNodes, nl.numGhostNodes) != (e.nInternal, e.nGhost):
            e.set_nodes(nl.numInternalNodes, nl.numGhostNodes)       # a changed node count invalidates the connectivity
        todo = {}
        for abi in names:
            key = _key(STATE_KEYS[abi], nl.name)
            if key in state:
                todo[abi] = state[key]
        if todo:
            e.upload_state(_keep_connectivity=keep_connectivity, **todo)

    def markDirty(self, *abiNames):
        """Kept for source compatibility with round 1: every call of this package now re-reads the State's fields."""
        return None

    def _require_connectivity(self, dataBase, who):
        if not (self._engine.connectivity_valid() and dataBase.connectivityValid):
            raise SPHB200Error("SPH::%s: no valid connectivity (call updateConnectivity after the node set or the ghost nodes changed)" % who)

    def evaluateDerivatives(self, time, dt, dataBase, state, derivs):
        """SPH<Dim>::evaluateDerivatives (SPH.cc:141-555) followed by the post sub-package's evaluateDerivatives
        (Integrator.cc:217-229 loops the packages); results land in the enrolled derivative fields."""
        if self._engine is None:
            raise SPHB200Error("SPH: initializeProblemStartup was not called")
        nl = dataBase.nodeLists[0]
        needed = ["position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy", "pressure", "soundSpeed", "omegaGradh"]
        for abi in ("DvDxQ", "fCl", "fCq"):
            if state.registered(STATE_KEYS[abi], nl.name):
                needed.append(abi)
        self._sync_state(nl, state, needed)
        self._require_connectivity(dataBase, "evaluateDerivatives")
        got = self._engine.evaluate_derivatives_to_host(time=time, dt=dt)      # pair loop and download overlapped, same bits
        for abi, key in DERIV_KEYS.items():
            k = _key(key, nl.name)
            if k in derivs:
                derivs[k][...] = got[abi].reshape(derivs[k].shape)

    def preStepInitialize(self, dataBase, state, derivs):
        """SPHBase::preStepInitialize (SPH/SPHBase.cc:322-352): with RigorousSumDensity the mass density is replaced by the SPH
        sum (computeSPHSumMassDensity) before the step; done on the device, the State's field is refreshed."""
        if self.densityUpdate != RigorousSumDensity or self._engine is None:
            return
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("position", "H", "mass", "massDensity"))
        self._require_connectivity(dataBase, "preStepInitialize")
        self._engine.sum_mass_density()
        k = _key(HydroFieldNames.massDensity, nl.name)
        state[k][...] = self._engine.download_state("massDensity")["massDensity"]

    def postStateUpdate(self, time, dt, dataBase, state, derivs):
        """ArtificialViscosityHandle::postStateUpdate copies DvDx into the Q's velocity gradient
        (ArtificialViscosityHandle.cc:165-180) and SPHBase::postStateUpdate recomputes the grad-h corrections
        (SPHBase.cc:527-547, computeSPHOmegaGradhCorrection); both device-side.  Returns True when boundaries must be
        re-applied (the grad-h field changed), as the reference does."""
        if self._engine is None:
            return False
        nl = dataBase.nodeLists[0]
        if HydroFieldNames.ArtificialViscosityVelocityGradient in self._own:
            self._engine.copy_DvDx_to_Q()
            k = _key(HydroFieldNames.ArtificialViscosityVelocityGradient, nl.name)
            if k in state:                                # the State's copy follows the device (every later call re-reads the State)
                state[k][...] = self._engine.download_state("DvDxQ")["DvDxQ"].reshape(state[k].shape)
        if not self.gradhCorrection:
            return False
        # the reference evaluates the corrections on the connectivity of the step start although positions moved
        # (CheapSynchronousRK2.cc:76-84): the refresh of positions and H keeps the device's pair lists
        self._sync_state(nl, state, ("position", "H"))
        self._require_connectivity(dataBase, "postStateUpdate")
        self._engine.compute_omega_gradh()
        k = _key(HydroFieldNames.omegaGradh, nl.name)
        if k in state:
            state[k][...] = self._engine.download_state("omegaGradh")["omegaGradh"]
        return True

    def dt(self, dataBase, state, derivs, currentTime=0.0):
        """GenericHydro::dt (Physics/GenericHydro.cc:112-381) -> (dt, reason), from the state on the host and the derivatives
        of the last evaluateDerivatives call on the device."""
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("velocity", "H", "massDensity", "soundSpeed"))
        vote, why, node = self._engine.compute_dt(self.cfl, self.useVelocityMagnitudeForDt)
        return vote, "%s limit: dt = %g (node %d)" % (why.capitalize(), vote, node)

    def updateSpecificThermalEnergy(self, multiplier, dataBase, state, derivs):
        """SpecificThermalEnergyPolicy::update (Hydro/SpecificThermalEnergyPolicy.cc:47-174) on the device; the new eps is
        copied back into the State's field."""
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("velocity", "mass", "specificThermalEnergy"))
        self._engine.update_energy_compatible(multiplier)
        eps = self._engine.download_state("specificThermalEnergy")["specificThermalEnergy"]
        k = _key(HydroFieldNames.specificThermalEnergy, nl.name)
        state[k][...] = eps


class CRKSPHB200(SPHB200):
    """Drop-in for CRKSPH<Dim> (CRKSPH/CRKSPH.hh, CRKSPHBase.hh) together with the RKCorrections package the controller
    inserts in front of it (SpheralController.py:690-745; RK/RKCorrections.cc), for RKOrder.LinearOrder and RKSumVolume --
    the settings of tests/functional/Hydro/Sedov/Sedov-spherical-3d.py:60-61.  Hooks, in the integrator's call order:
      preStepInitialize  RKCorrections::preStepInitialize (volumes, RKCorrections.cc:298-340) then
                         CRKSPHBase::preStepInitialize (RigorousSumDensity, CRKSPHBase.cc:229-256)
      initialize         RKCorrections::initialize (corrections, RKCorrections.cc:346-372); returns True so that the
                         integrator re-applies ghost boundaries (to the corrections)
      evaluateDerivatives CRKSPH::evaluateDerivatives (CRKSPH.cc:148-440) + smoothing-scale sub-package."""
    _hydro = L.HYDRO_CRKSPH

    def __init__(self, dataBase, Q, W, order=RKOrder.LinearOrder, cfl=0.25, useVelocityMagnitudeForDt=False,
                 compatibleEnergyEvolution=True, evolveTotalEnergy=False, XSPH=True, densityUpdate=RigorousSumDensity,
                 epsTensile=0.0, nTensile=4.0, volumeType=RKSumVolume, device=0):
        if order != RKOrder.LinearOrder:
            raise SPHB200Error("CRKSPH: only RKOrder.LinearOrder is implemented on the device path")
        if volumeType != RKSumVolume:
            raise SPHB200Error("CRKSPH: only RKVolumeType.RKSumVolume is implemented on the device path "
                               "(RKVoronoiVolume needs the polytope tessellation, out of scope)")
        super().__init__(dataBase=dataBase, Q=Q, W=W, WPi=W, cfl=cfl, useVelocityMagnitudeForDt=useVelocityMagnitudeForDt,
                         compatibleEnergyEvolution=compatibleEnergyEvolution, evolveTotalEnergy=evolveTotalEnergy,
                         gradhCorrection=False, XSPH=XSPH, correctVelocityGradient=False, densityUpdate=densityUpdate,
                         epsTensile=epsTensile, nTensile=nTensile, device=device)
        self._order, self.volumeType = order, volumeType

    def label(self): return "CRKSPH"
    correctionOrder = property(lambda s: s._order)
    def requireReproducingKernels(self): return {RKOrder.ZerothOrder, self._order}

    def _resize_owned(self, nl):
        super()._resize_owned(nl)
        n, nd = nl.numNodes, self._ndim
        own = self._own
        if HydroFieldNames.volume not in own or own[HydroFieldNames.volume].shape[0] != n:
            own[HydroFieldNames.volume] = np.zeros(n)
        k = RKFieldNames.rkCorrections(self._order)
        if k not in own or own[k].shape[0] != n:
            own[k] = np.zeros((n, (nd + 1)*(nd + 1)))
            own[k][:, 0] = 1.0

    def registerState(self, dataBase, state):
        """RKCorrections::registerState (volume, corrections; RKCorrections.cc:150-176) + CRKSPHBase::registerState
        (CRKSPHBase.cc:131-187)."""
        super().registerState(dataBase, state)
        nl = dataBase.nodeLists[0]
        state.enroll(nl.name, HydroFieldNames.volume, self._own[HydroFieldNames.volume])
        k = RKFieldNames.rkCorrections(self._order)
        state.enroll(nl.name, k, self._own[k])

    def _pull(self, nl, state, abi):
        """device -> the State's host array for one field (internal values were produced on the device)."""
        got = self._engine.download_state(abi)[abi]
        k = _key(STATE_KEYS[abi], nl.name)
        state[k][...] = got.reshape(state[k].shape)

    def preStepInitialize(self, dataBase, state, derivs):
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("position", "H", "mass", "volume"))
        self._require_connectivity(dataBase, "preStepInitialize")
        self._engine.crk_compute_volume()
        self._pull(nl, state, "volume")
        if self.densityUpdate == RigorousSumDensity:
            self._engine.crk_sum_mass_density(nl.rhoMin, nl.rhoMax)
            self._pull(nl, state, "massDensity")

    def initialize(self, time, dt, dataBase, state, derivs):
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("volume", "rkCorrections"))      # ghost values set by the host's boundary conditions
        self._engine.crk_compute_corrections()
        self._pull(nl, state, "rkCorrections")
        return True

    def evaluateDerivatives(self, time, dt, dataBase, state, derivs):
        nl = dataBase.nodeLists[0]
        self._sync_state(nl, state, ("volume", "rkCorrections"))
        super().evaluateDerivatives(time, dt, dataBase, state, derivs)


# ---- factories (SPH/SPHHydros.py:9-140, :145-208) ----------------------------------------------------------------------------
def SPH(W, WPi=None, WGrad=None, dataBase=None, Q=None, filter=None, cfl=0.25, useVelocityMagnitudeForDt=False,
        compatibleEnergyEvolution=True, evolveTotalEnergy=False, gradhCorrection=True, XSPH=True,
        correctVelocityGradient=True, sumMassDensityOverAllNodeLists=True, densityUpdate=RigorousSumDensity,
        HUpdate=IdealH, epsTensile=0.0, nTensile=4.0, damageRelieveRubble=False, strengthInDamage=False,
        xmin=(-1e100, -1e100, -1e100), xmax=(1e100, 1e100, 1e100), etaMinAxis=0.1, ASPH=False,
        smoothingScaleMethod=None, device=0):
    if dataBase is None:
        raise SPHB200Error("SPH: dataBase is required")
    if dataBase.numSolidNodeLists > 0:
        raise RuntimeError("Cannot mix solid and fluid NodeLists.")
    if WPi is None:
        WPi = W
    if not Q:                                                     # SPHHydros.py:88-94
        Cl = 2.0*(W.kernelExtent/2.0)
        Cq = 2.0*(W.kernelExtent/2.0)**2
        Q = LimitedMonaghanGingoldViscosity(Clinear=Cl, Cquadratic=Cq, kernel=WPi)
    result = SPHB200(dataBase=dataBase, Q=Q, W=W, WPi=WPi, cfl=cfl, useVelocityMagnitudeForDt=useVelocityMagnitudeForDt,
                     compatibleEnergyEvolution=compatibleEnergyEvolution, evolveTotalEnergy=evolveTotalEnergy,
                     gradhCorrection=gradhCorrection, XSPH=XSPH, correctVelocityGradient=correctVelocityGradient,
                     sumMassDensityOverAllNodeLists=sumMassDensityOverAllNodeLists, densityUpdate=densityUpdate,
                     epsTensile=epsTensile, nTensile=nTensile, xmin=xmin, xmax=xmax, device=device)
    result.prependSubPackage(Q)                                   # SPHHydros.py:125-127
    if smoothingScaleMethod is None:                              # SPHHydros.py:129-140
        if isinstance(ASPH, str) and ASPH.upper() == "CLASSIC":
            smoothingScaleMethod = ASPHClassicSmoothingScale(HUpdate, W)
        else:
            smoothingScaleMethod = ASPHSmoothingScale(HUpdate, W) if ASPH else SPHSmoothingScale(HUpdate, W)
    result._smoothingScaleMethod = smoothingScaleMethod
    result.appendSubPackage(smoothingScaleMethod)
    return result


def ASPH(W, **kw):
    """SPHHydros.py:145-208"""
    kw["ASPH"] = True
    return SPH(W, **kw)


def CRKSPH(dataBase, W, Q=None, order=RKOrder.LinearOrder, filter=0.0, cfl=0.25, useVelocityMagnitudeForDt=False,
           compatibleEnergyEvolution=True, evolveTotalEnergy=False, XSPH=True, densityUpdate=RigorousSumDensity,
           HUpdate=IdealH, epsTensile=0.0, nTensile=4.0, damageRelieveRubble=False, ASPH=False, etaMinAxis=0.1,
           crktype="default", smoothingScaleMethod=None, volumeType=RKSumVolume, device=0):
    """CRKSPH/CRKSPHHydros.py:9-112 -- same keywords and defaults (volumeType is the controller's keyword,
    SpheralController.py:53; here the device path implements RKSumVolume)."""
    if dataBase.numSolidNodeLists > 0:
        raise RuntimeError("Cannot mix solid and fluid NodeLists.")
    if crktype.lower() != "default":
        raise SPHB200Error("CRKSPH: only crktype='default' is implemented on the device path")
    if not Q:                                                     # CRKSPHHydros.py:64-68
        Cl = 2.0*(W.kernelExtent/4.0)
        Cq = 1.0*(W.kernelExtent/4.0)**2
        Q = LimitedMonaghanGingoldViscosity(Clinear=Cl, Cquadratic=Cq, kernel=W)
    result = CRKSPHB200(dataBase=dataBase, Q=Q, W=W, order=order, cfl=cfl, useVelocityMagnitudeForDt=useVelocityMagnitudeForDt,
                        compatibleEnergyEvolution=compatibleEnergyEvolution, evolveTotalEnergy=evolveTotalEnergy, XSPH=XSPH,
                        densityUpdate=densityUpdate, epsTensile=epsTensile, nTensile=nTensile, volumeType=volumeType,
                        device=device)
    result.prependSubPackage(Q)                                   # CRKSPHHydros.py:91
    if smoothingScaleMethod is None:                              # CRKSPHHydros.py:94-103
        if isinstance(ASPH, str) and ASPH.upper() == "CLASSIC":
            smoothingScaleMethod = ASPHClassicSmoothingScale(HUpdate, W)
        else:
            smoothingScaleMethod = ASPHSmoothingScale(HUpdate, W) if ASPH else SPHSmoothingScale(HUpdate, W)
    result._smoothingScaleMethod = smoothingScaleMethod
    result.appendSubPackage(smoothingScaleMethod)
    return result


def ACRKSPH(*args, **kw):
    """CRKSPHHydros.py:115-117"""
    kw["ASPH"] = True
    return CRKSPH(*args, **kw)
