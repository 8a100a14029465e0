/* ============================================================================
 * sphb200.h -- C ABI of the B200-native SPH hydro-derivative engine.
 *
 * This is the drop-in boundary: a host-side Spheral `Physics<Dim>` package (or the Python
 * mirror in spheral_b200/) calls these entry points instead of running the reference's CPU loops.
 * Plain pointers and sizes only; no C++/torch types.  Every entry point cites the reference interface it
 * replaces (llnl/spheral @ c64796e1, paths relative to src/).
 *
 * Conventions
 *   - One context per GPU / rank.  Not re-entrant per context.  All work is stream-ordered on the context's
 *     stream; functions that return data to the host synchronise that stream.
 *   - All functions return 0 on success, non-zero on error; sphb200_last_error() gives the message
 *     (the reference throws VERIFYError -> Python exception, Utilities/DBC.hh:28-45; the host wrapper re-raises).
 *   - Host field layout is the reference's Field<DataType> AoS (Field/Field.hh:213):
 *       Vector ndim doubles | SymTensor 3-D xx,xy,xz,yy,yz,zz / 2-D xx,xy,yy | Tensor ndim*ndim row-major,
 *     nodes [0,nInternal) internal then [nInternal,nInternal+nGhost) ghost (FieldView.hh:64-66).
 *   - Derivatives are defined on internal nodes only (the reference's thread reduction covers internal nodes
 *     only, Utilities/OpenMP_wrapper.hh:69-117); ghost entries are written as 0.
 * ==========================================================================*/
#ifndef SPHB200_H
#define SPHB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB200_ABI_VERSION 5

typedef struct sphb200_ctx sphb200_ctx;

/* ArtificialViscosity flavour (ArtificialViscosity/MonaghanGingoldViscosity.cc:41-101,
   LimitedMonaghanGingoldViscosity.cc:120-219) */
enum { SPHB200_Q_MG = 0, SPHB200_Q_LIMITED_MG = 1 };
/* smoothing-scale sub-package run after the hydro (SPH/SPHHydros.py:129-140):
   SPHSmoothingScale.cc:101-275 | ASPHSmoothingScale.cc:110-147 | none */
enum { SPHB200_H_SPH = 0, SPHB200_H_ASPH = 1, SPHB200_H_NONE = 2,
       SPHB200_H_ASPH_CLASSIC = 3 /* ASPHClassicSmoothingScale (SmoothingScale/ASPHClassicSmoothingScale.cc; SPHHydros.py:133-134, ASPH = "Classic"):
                                     the ASPH tensor derivative plus the second-moment ideal H, behind either hydro */ };
/* analytic kernels for sphb200_table_kernel_build (Kernel/<name>KernelInline.hh) */
enum { SPHB200_KERNEL_BSPLINE = 0, SPHB200_KERNEL_WENDLANDC4 = 1, SPHB200_KERNEL_WENDLANDC2 = 2,
       SPHB200_KERNEL_NBSPLINE = 100 /* + order (1..11): NBSplineKernel(order), Kernel/NBSplineKernel.cc:17-122 -- the kernel of the stock
                                        Noh scripts (Noh-planar-1d.py, Noh-spherical-3d.py: NBSplineKernel(5)) */ };
/* hydro flavour served by the context: SPH<Dim> (SPH/SPH.cc) | CRKSPH<Dim> (CRKSPH/CRKSPH.cc; RKOrder::LinearOrder,
   RKVolumeType::RKSumVolume -- the settings of tests/functional/Hydro/Sedov/Sedov-spherical-3d.py:60-61) */
enum { SPHB200_HYDRO_SPH = 0, SPHB200_HYDRO_CRKSPH = 1 };
/* which TableKernel a table upload refers to (SPHBase.hh:182-183: kernel() / PiKernel()) */
enum { SPHB200_TABLE_W = 0, SPHB200_TABLE_WPI = 1 };

/* Constructor arguments of SPH<Dim> (SPH/SPH.hh:40-56, SPH/SPHHydros.py:9-33) that the derivative path
   reads, plus the Q parameters (ArtificialViscosityHandle.cc:37-53) and NodeList bounds used by the
   smoothing-scale package (NodeList hmin/hmax/nodesPerSmoothingScale). */
typedef struct {
  int    ndim;                      /* 2 | 3 */
  int    compatibleEnergy;          /* compatibleEnergyEvolution */
  int    evolveTotalEnergy;
  int    XSPH;
  int    correctVelocityGradient;
  double epsTensile;
  double nTensile;
  double nPerh;                     /* nodesPerSmoothingScale of NodeList 0 */
  int    Qkind;                     /* SPHB200_Q_* */
  double Cl, Cq, eps2, negligibleSoundSpeed;
  int    balsara, linearInExpansion, quadraticInExpansion;
  double etaCritFrac, etaFoldFrac;
  int    hEvolution;                /* SPHB200_H_* */
  double hmin, hmax;
  int    hydro;                     /* SPHB200_HYDRO_* (0 = SPH) */
  double hminratio;                 /* NodeList::hminratio; read by SPHB200_H_ASPH_CLASSIC only (ASPHClassicSmoothingScale.cc:254, 302, 372) */
} sphb200_options;

/* State fields read by the path (SPH.cc:206-216; Appendix B of SURVEY.md). Bits for fieldMask. */
enum {
  SPHB200_F_POSITION = 1u << 0,  SPHB200_F_VELOCITY = 1u << 1,  SPHB200_F_H = 1u << 2,
  SPHB200_F_MASS = 1u << 3,      SPHB200_F_RHO = 1u << 4,       SPHB200_F_EPS = 1u << 5,
  SPHB200_F_PRESSURE = 1u << 6,  SPHB200_F_SOUNDSPEED = 1u << 7, SPHB200_F_OMEGA = 1u << 8,
  SPHB200_F_DVDXQ = 1u << 9,     SPHB200_F_FCL = 1u << 10,      SPHB200_F_FCQ = 1u << 11,
  /* CRKSPH only: HydroFieldNames::volume (Scalar) and RKFieldNames::rkCorrections(LinearOrder)
     ((1+ndim)^2 doubles per node: C[1+ndim], then dC_d[1+ndim] per direction; RK/RKUtilitiesInline.hh:58-105) */
  SPHB200_F_VOLUME = 1u << 12,   SPHB200_F_RKCORR = 1u << 13,
  SPHB200_F_ALL_BASIC = 0x1FFu
};
typedef struct {
  const double *position, *velocity, *H, *mass, *massDensity, *specificThermalEnergy,
               *pressure, *soundSpeed, *omegaGradh, *DvDxQ, *fCl, *fCq, *volume, *rkCorrections;
} sphb200_host_state;

/* Derivative fields written by the path (SPH.cc:230-245, SPHSmoothingScale.cc:130-137). Bits for fieldMask. */
enum {
  SPHB200_D_DXDT = 1u << 0,   SPHB200_D_DRHODT = 1u << 1,  SPHB200_D_DVDT = 1u << 2,   SPHB200_D_DEPSDT = 1u << 3,
  SPHB200_D_DVDX = 1u << 4,   SPHB200_D_LOCALDVDX = 1u << 5, SPHB200_D_GRADRHO = 1u << 6, SPHB200_D_M = 1u << 7,
  SPHB200_D_LOCALM = 1u << 8, SPHB200_D_RHOSUM = 1u << 9,  SPHB200_D_NORM = 1u << 10,  SPHB200_D_MAXQ = 1u << 11,
  SPHB200_D_EFFQ = 1u << 12,  SPHB200_D_XSPHW = 1u << 13,  SPHB200_D_XSPHDV = 1u << 14, SPHB200_D_DHDT = 1u << 15,
  SPHB200_D_HIDEAL = 1u << 16, SPHB200_D_M0 = 1u << 17,    SPHB200_D_M1 = 1u << 18,
  SPHB200_D_ALL = 0x7FFFFu
};
typedef struct {
  double *DxDt, *DrhoDt, *DvDt, *DepsDt, *DvDx, *localDvDx, *gradRho, *M, *localM,
         *rhoSum, *normalization, *maxViscousPressure, *effViscousPressure, *XSPHWeightSum, *XSPHDeltaV,
         *DHDt, *Hideal, *massZerothMoment, *massFirstMoment;
} sphb200_host_derivs;

/* ---- life cycle --------------------------------------------------------------------------------------
   replaces: SPH<Dim>::SPH(...) (SPH/SPH.cc:43-80) + SPHBase::initializeProblemStartup (SPHBase.cc:116-142) */
int  sphb200_abi_version(void);
int  sphb200_create(sphb200_ctx** ctx, int device, const sphb200_options* opts);
void sphb200_destroy(sphb200_ctx* ctx);
int  sphb200_set_options(sphb200_ctx* ctx, const sphb200_options* opts);   /* property setters, SPHBase.hh:131-176 */
const char* sphb200_last_error(const sphb200_ctx* ctx);                     /* NULL ctx -> last global error */
int  sphb200_sync(sphb200_ctx* ctx);

/* ---- TableKernel ---------------------------------------------------------------------------------------
   Upload the coefficients of a host TableKernel: three QuadraticInterpolators sharing (xmin,xstep,n1), each
   3*(n1+1) doubles (Kernel/TableKernel.cc:169-209; readable from a real Spheral through
   PYB11/Kernel/Kernel.py:432-434 + PYB11/Utilities/QuadraticInterpolator.py:53-102).  nperhVals/wsumVals are the
   CubicHermite lookups (n values then n gradients; may be NULL when hEvolution != SPHB200_H_SPH). */
int  sphb200_set_kernel_table(sphb200_ctx* ctx, int which, double kext, double xmin, double xstep, size_t n1,
                              const double* Wcoef, const double* gradWcoef, const double* grad2Wcoef,
                              size_t nperhN, double nperhXmin, double nperhXmax, const double* nperhVals,
                              size_t wsumN, double wsumXmin, double wsumXmax, const double* wsumVals);
/* Stand-alone host-side TableKernel construction (no GPU needed; used when no Spheral TableKernel exists):
   TableKernel<Dim>::TableKernel(kernel, numPoints, minNperh, maxNperh), Kernel/TableKernel.cc:169-209.
   Coefficient arrays must hold sphb200_table_ncoef(numPoints) doubles, lookup arrays 2*numPoints doubles. */
size_t sphb200_table_ncoef(size_t numPoints);
int  sphb200_table_kernel_build(int kind, int ndim, size_t numPoints, double minNperh, double maxNperh,
                                double* kext, double* xstep, size_t* n1,
                                double* Wcoef, double* gradWcoef, double* grad2Wcoef,
                                double* nperhVals, double* nperhRange /*[2]*/,
                                double* wsumVals, double* wsumRange /*[2]*/);

/* ---- node data -------------------------------------------------------------------------------------------
   replaces: State<Dim>::fields(name) reads at SPH.cc:206-216 (device-resident copies keep their values across
   Integrator stages until re-uploaded). */
int  sphb200_set_nodes(sphb200_ctx* ctx, size_t nInternal, size_t nGhost);
int  sphb200_upload_state(sphb200_ctx* ctx, unsigned fieldMask, const sphb200_host_state* s);
/* The same copy, but the pair lists on the device stay valid even if positions / H are among the fields: a package evaluated
   in the middle of a step works on the ConnectivityMap of the step start although the nodes have moved since
   (CheapSynchronousRK2.cc:76-99: state.update, then evaluateDerivatives, no neighbour update in between).
   replaces: the State<Dim>::fields(name) reads of a mid-step evaluateDerivatives / postStateUpdate. */
int  sphb200_upload_state_values(sphb200_ctx* ctx, unsigned fieldMask, const sphb200_host_state* s);
int  sphb200_download_state(sphb200_ctx* ctx, unsigned fieldMask, double* const* fields /* same order as struct */);

/* ---- neighbour pairs ---------------------------------------------------------------------------------------
   replaces: Neighbor::updateNodes (TreeNeighbor.cc:370-455 / NestedGridNeighbor.cc:209-560) +
             ConnectivityMap::computeConnectivity (ConnectivityMap.cc:747-1152).
   Uses the positions/H currently on the device.  npairs = size of the NodePairList (i<j once). */
int  sphb200_build_pairs(sphb200_ctx* ctx, size_t* npairs);
/* 1 if the pair lists on the device match the positions / H there (no upload or halo landing of either since the last build);
   the role of DataBase::connectivityMap() being current.  No device work. */
int  sphb200_connectivity_valid(const sphb200_ctx* ctx);
/* Mask (SPHB200_STATE_* bits) of the state fields that currently hold values on the device, whether uploaded, received as ghosts or
   produced there (sphb200_copy_DvDx_to_Q, the CRKSPH volumes ...).  The boundary conditions apply to every field REGISTERED in the
   State (Integrator::applyGhostBoundaries, Integrator.cc:530-570; ArtificialViscosityHandle.cc:105-119 registers the Q's velocity
   gradient and Cl / Cq multipliers): a halo exchange asks here which fields must travel.  No device work. */
unsigned sphb200_state_fields_present(const sphb200_ctx* ctx);
/* NodePairList in the reference order (NodePairIdxType::operator<, NodePairIdxType.hh:34-58): sorted (i,j), i<j. */
int  sphb200_download_pairs(sphb200_ctx* ctx, uint32_t* i, uint32_t* j, size_t cap);
/* ConnectivityMap::numNeighborsForNode for internal nodes (ConnectivityMapInline.hh) */
int  sphb200_download_neighbor_counts(sphb200_ctx* ctx, uint32_t* counts);

/* ---- derivatives ---------------------------------------------------------------------------------------------
   replaces: SPH<Dim>::evaluateDerivatives (SPH.cc:141-555) followed by the smoothing-scale package's
   evaluateDerivatives; derivatives are zeroed first (CheapSynchronousRK2.cc:87 derivs.Zero()). Asynchronous. */
int  sphb200_evaluate_derivatives(sphb200_ctx* ctx, double time, double dt);
int  sphb200_download_derivs(sphb200_ctx* ctx, unsigned fieldMask, const sphb200_host_derivs* d);
/* evaluateDerivatives as the reference's caller sees it -- SPH<Dim>::evaluateDerivatives(time, dt, dataBase, state, derivs) fills
   HOST fields (SPH.cc:141-163; Integrator.cc:217-229 evaluateDerivatives) -- in one call: sphb200_evaluate_derivatives followed by
   sphb200_download_derivs of fieldMask, with the two overlapped.  The pair loop runs in chunks of the HOST index range (4 by default,
   SPHB200_E2H_CHUNKS; each chunk = the tiles that hold its nodes); while chunk q+1 is computed, chunk q's slice of every selected
   field is un-permuted and copied to the host on the copy stream.  Results are bit-identical to the two separate calls (a node's
   sums do not depend on the launch that forms them).  Falls back to the two calls when the host order has no spatial coherence (a
   tile would be visited by every chunk), below 256 k internal nodes, and for CRKSPH.  Synchronises (the host fields are complete on
   return); the derivatives also stay on the device as after sphb200_evaluate_derivatives. */
int  sphb200_evaluate_derivatives_to_host(sphb200_ctx* ctx, double time, double dt, unsigned fieldMask, const sphb200_host_derivs* d);
/* Restart: SPHBase::restoreState (SPH/SPHBase.cc:741-765) reads the package-owned derivative fields back, because
   CheapSynchronousRK2 advances the first trial state after a restart with them.  Uploads node-wise derivative fields (host AoS,
   original order); sphb200_state_update / sphb200_compute_dt / sphb200_copy_DvDx_to_Q then work as after an evaluation.  Pair-wise
   data are not restored: the compatible energy update needs a new sphb200_evaluate_derivatives, as in the reference, where the
   PairwiseField is rebuilt with the connectivity (SPH.cc:129-133). */
int  sphb200_upload_derivs(sphb200_ctx* ctx, unsigned fieldMask, const sphb200_host_derivs* d);
/* PairwiseField "pair-wise accelerations" (SPH.cc:129-133, :430) in NodePairList order, npairs*ndim doubles. */
int  sphb200_download_pair_accelerations(sphb200_ctx* ctx, double* pairAccelerations, size_t cap);
/* ArtificialViscosityHandle::postStateUpdate copy DvDx -> Q velocity gradient (ArtificialViscosityHandle.cc:165-180) */
int  sphb200_copy_DvDx_to_Q(sphb200_ctx* ctx);

/* ---- CRKSPH (contexts created with hydro = SPHB200_HYDRO_CRKSPH) --------------------------------
   The reference sequence per stage is RKCorrections::preStepInitialize (volumes, RK/RKCorrections.cc:298-340) ->
   boundaries -> RKCorrections::initialize (corrections, :346-372) -> boundaries -> CRKSPH::evaluateDerivatives.
   Each call below writes the INTERNAL entries of its field on the device; ghost entries are the caller's (boundary
   conditions / halo exchange), exactly as in the reference, and travel with upload_state / halo_pack+unpack
   under SPHB200_F_VOLUME / SPHB200_F_RKCORR.
     sphb200_crk_compute_volume       computeRKSumVolume                   RK/computeRKSumVolume.cc:33-116
     sphb200_crk_compute_corrections  RKUtilities::computeCorrections      RK/RKUtilities.cc:252-491 (LinearOrder)
     sphb200_crk_sum_mass_density     computeCRKSPHSumMassDensity          CRKSPH/computeCRKSPHSumMassDensity.cc:21-131
   sphb200_evaluate_derivatives then runs CRKSPH<Dim>::evaluateDerivativesImpl (CRKSPH/CRKSPH.cc:176-440) + the
   smoothing-scale sub-package; pair accelerations and sphb200_update_energy_compatible work as for SPH. */
int  sphb200_crk_compute_volume(sphb200_ctx* ctx);
int  sphb200_crk_compute_corrections(sphb200_ctx* ctx);
int  sphb200_crk_sum_mass_density(sphb200_ctx* ctx, double rhoMin, double rhoMax);

/* ---- per-step callers of the derivative path, device-resident ----------------------------------
   With these the state never leaves the GPU between the stages of CheapSynchronousRK2 (Integrator/CheapSynchronousRK2.cc:40-132);
   the host integrator (spheral_b200/integrator.py, or a C++ Integrator subclass) reads back one number per step: dt.
     sphb200_sum_mass_density      computeSPHSumMassDensity (SPHBase::preStepInitialize, RigorousSumDensity)
                                   SPH/computeSPHSumMassDensity.cc:15-89, SPH/SPHBase.cc:336-352.  Internal rho entries are written.
     sphb200_compute_omega_gradh   computeSPHOmegaGradhCorrection (SPHBase::postStateUpdate)
                                   SPH/computeSPHOmegaGradhCorrection.cc:20-113, SPH/SPHBase.cc:539-547
     sphb200_update_eos_gamma_law  PressurePolicy / SoundSpeedPolicy with GammaLawGas: P and cs of every node
                                   Hydro/PressurePolicy.cc, Material/GammaLawGas.cc:185-189, 233-238, EquationOfStateInline.hh:98-106
     sphb200_state_copy / _assign  State::copyState / State::assign (CheapSynchronousRK2.cc:70-71, 104, 113)
     sphb200_state_update          State::update(derivs, multiplier, t, dt) (DataBase/State.cc:221-300) for the policies the hydro and
                                   smoothing-scale packages register (SPHBase.cc:203-261, SPH.cc:96-113, SmoothingScaleBase.cc:56-83,
                                   ASPHSmoothingScale.cc:72-90): rho IncrementBoundedState, position / velocity IncrementState,
                                   eps SpecificThermalEnergyPolicy (compatible) or IncrementState, H IncrementBoundedState /
                                   ReplaceBoundedState / IncrementASPHHtensor, then pressure and sound speed.  timeAdvanceOnly != 0
                                   degrades every policy to its increment (updateAsIncrement).  Uses the derivatives of the last
                                   sphb200_evaluate_derivatives call, which stay usable after a later sphb200_build_pairs.
     sphb200_compute_dt            GenericHydro::dt (Physics/GenericHydro.cc:112-381); reason: 0 sound speed, 1 artificial
                                   viscosity, 2 velocity divergence, 3 acceleration, 4 velocity magnitude, 5 pairwise velocity
                                   difference; node = the limiting node.  Synchronises (one 16-byte read-back). */
typedef struct {
  double gamma;                               /* GammaLawGas */
  double minimumPressure, maximumPressure, externalPressure;
  int    minPressureType;                     /* 0 PressureFloor | 1 ZeroPressure */
} sphb200_gamma_law;
enum { SPHB200_HEVOLUTION_IDEALH = 0, SPHB200_HEVOLUTION_INTEGRATEH = 1, SPHB200_HEVOLUTION_FIXEDH = 2 };
typedef struct {
  sphb200_gamma_law eos;
  double rhoMin, rhoMax;                      /* FluidNodeList::rhoMin / rhoMax */
  double hminratio;                           /* NodeList::hminratio (ASPH) */
  int    HEvolution;                          /* SPHB200_HEVOLUTION_* (SmoothingScaleBase.hh) */
} sphb200_step_options;
int  sphb200_sum_mass_density(sphb200_ctx* ctx);
int  sphb200_compute_omega_gradh(sphb200_ctx* ctx);
int  sphb200_update_eos_gamma_law(sphb200_ctx* ctx, const sphb200_gamma_law* eos);
int  sphb200_state_copy(sphb200_ctx* ctx);
int  sphb200_state_assign(sphb200_ctx* ctx);
int  sphb200_state_update(sphb200_ctx* ctx, const sphb200_step_options* so, double multiplier, int timeAdvanceOnly);
/* One sweep of iterateIdealH (Utilities/iterateIdealH.cc:120-190), the start-up relaxation of H: every internal node that has
   not converged yet takes the "new H" of the last sphb200_evaluate_derivatives call; maxDeltaH = max over those nodes of
   max|phi - 1|, phi the eigenvalues of H1^(1/2) H^-1 H1^(1/2); nodes with deltaH <= tolerance are frozen.  The caller loops
   build_pairs -> evaluate_derivatives -> iterate_ideal_h until maxDeltaH <= tolerance (firstSweep != 0 clears the frozen set).
   SPH smoothing scale (isotropic ideal H: phi = h1/lambda(H)) and the classic ASPH one (tensor ideal H: the general formula). */
int  sphb200_iterate_ideal_h(sphb200_ctx* ctx, int firstSweep, double tolerance, double* maxDeltaH);
int  sphb200_compute_dt(sphb200_ctx* ctx, double cfl, int useVelocityMagnitudeForDt, double* dt, int* reason, uint32_t* node);

/* ---- reflecting-plane boundaries on the device -------------------------------------------
   planes: nPlanes points and inward normals, ndim doubles each (the stock Noh / Sedov octant: the coordinate planes).
     sphb200_reflect_set_ghost_nodes   PlanarBoundary::setGhostNodes per plane in order (Integrator/Integrator.cc:415-424,
                                       Boundary/findNodesTouchingThroughPlanes.cc, Boundary/mapPositionThroughPlanes.hh:17-27):
                                       discards the current ghosts, selects the control nodes of each plane on the device
                                       (ascending node order; later planes mirror the ghosts of earlier ones), resizes the node
                                       set and fills every field present.  Internal state is kept.  One host round trip per plane.
     sphb200_reflect_apply_ghosts      ReflectingBoundary::applyGhostBoundary (Boundary/ReflectingBoundary.cc:182-250) for the
                                       masked state fields: scalars copied, vectors R v, tensors R (T R), H (R (H R)).Symmetric().
                                       The connectivity is kept (the reference refreshes ghost values mid-step on a fixed one).
     sphb200_reflect_finalize_derivatives  SPHBase::finalizeDerivatives (SPH/SPHBase.cc:502-519): with compatible energy the ghost
                                       entries of the acceleration (R a) and of DepsDt (copy) are set from their control nodes --
                                       SpecificThermalEnergyPolicy reads them for the ghost end of an internal-ghost pair
     sphb200_reflect_enforce           PlanarBoundary::enforceBoundary (Boundary/PlanarBoundary.cc:153-190,
                                       ReflectingBoundary.cc:255-330): internal nodes behind a plane are mirrored back, their
                                       velocity reflected.  nViolations may be NULL (then no host synchronisation). */
int  sphb200_reflect_configure(sphb200_ctx* ctx, int nPlanes, const double* points, const double* normals);
/* General form: planar boundaries with an enter and an exit plane (Boundary/PlanarBoundary.hh).  Reflecting: exit == enter
   (exit arrays ignored).  Periodic (Boundary/PeriodicBoundary.cc:60-62): a PeriodicBoundary(plane1, plane2) is TWO entries,
   (enter = plane1, exit = plane2) and (enter = plane2, exit = plane1); ghosts are unreflected copies displaced through the
   planes (mapPositionThroughPlanes.hh:17-27), escaped nodes re-enter through the other plane.  The sphb200_reflect_* calls
   below serve both kinds. */
enum { SPHB200_BOUNDARY_REFLECTING = 0, SPHB200_BOUNDARY_PERIODIC = 1 };
int  sphb200_boundary_configure(sphb200_ctx* ctx, int nBoundaries, const int* kinds, const double* enterPoints, const double* enterNormals,
                                const double* exitPoints, const double* exitNormals);
int  sphb200_reflect_set_ghost_nodes(sphb200_ctx* ctx, size_t* nGhost);
int  sphb200_reflect_apply_ghosts(sphb200_ctx* ctx, unsigned fieldMask);
int  sphb200_reflect_finalize_derivatives(sphb200_ctx* ctx);
int  sphb200_reflect_enforce(sphb200_ctx* ctx, size_t* nViolations);

/* ---- compatible energy -----------------------------------------------------------------------------------------
   replaces: SpecificThermalEnergyPolicy::update (Hydro/SpecificThermalEnergyPolicy.cc:47-174):
   eps += multiplier * (pair-wise discrete work), using the velocity/mass/eps currently on the device and the
   DvDt, DepsDt and pair accelerations of the last sphb200_evaluate_derivatives call. */
int  sphb200_update_energy_compatible(sphb200_ctx* ctx, double multiplier);

/* ---- multi-GPU halo (Distributed/DistributedBoundary.cc:565-753, 1256-1353) ---------------------------------------
   The exchange itself is NCCL send/recv issued by the host plumbing on device staging buffers; these two calls
   are the device-side pack (gather the listed send nodes of the masked fields into one staging buffer) and the
   ghost landing (copy a received staging buffer into ghost slots [firstGhost, firstGhost+count)).
   Staging layout: field-major, each field count*width doubles, fields in ascending mask-bit order. */
size_t sphb200_halo_bytes_per_node(const sphb200_ctx* ctx, unsigned fieldMask);
int  sphb200_halo_pack(sphb200_ctx* ctx, unsigned fieldMask, const uint32_t* sendNodesDevice, size_t count,
                       void* stagingDevice);
int  sphb200_halo_unpack(sphb200_ctx* ctx, unsigned fieldMask, size_t firstGhost, size_t count,
                         const void* stagingDevice);
/* Ghost-value refresh on a FIXED connectivity (Integrator::applyGhostBoundaries between the stages of a step): like
   sphb200_halo_unpack, but positions / H landing in the ghost slots do not invalidate the pair lists. */
int  sphb200_halo_unpack_values(sphb200_ctx* ctx, unsigned fieldMask, size_t firstGhost, size_t count, const void* stagingDevice);
/* SPHBase::finalizeDerivatives across a domain boundary (SPH/SPHBase.cc:502-519; SURVEY 8e: "with compatible energy, one
   [exchange] after: DvDt, DepsDt on ghosts"): pack {DvDt (ndim), DepsDt} of the listed send nodes / land a received block in
   the ghost entries of the derivative arrays.  Staging: count*ndim doubles of DvDt, then count doubles of DepsDt. */
int  sphb200_halo_pack_derivs(sphb200_ctx* ctx, const uint32_t* sendNodesDevice, size_t count, void* stagingDevice);
int  sphb200_halo_unpack_derivs(sphb200_ctx* ctx, size_t firstGhost, size_t count, const void* stagingDevice);
/* Domain bounds for the ghost-set decision (the role of the bounding boxes all-gathered by
   NestedGridDistributedBoundary.cc:116-170 / TreeDistributedBoundary.cc:150-295): coordinate range and the largest
   per-axis kernel extent kext*sqrt((H^-2)_aa) (Neighbor::HExtent, NeighborInline.hh:52-64) over nodes [0,count). */
int  sphb200_node_bounds(sphb200_ctx* ctx, size_t count, double lo[3], double hi[3], double maxExtent[3]);
/* Send-node selection for a slab decomposition along `axis` (DistributedBoundary::buildSendNodes role): among nodes
   [0,count), those with x < lo + width go to the lower neighbour, those with x >= hi - width to the upper one.  Index
   lists are written in ascending node order (deterministic) into caller-provided device buffers of `cap` entries. */
int  sphb200_halo_select(sphb200_ctx* ctx, int axis, size_t count, double lo, double hi, double width,
                         uint32_t* sendLowDevice, size_t* nLow, uint32_t* sendHighDevice, size_t* nHigh, size_t cap);
/* Stream-ordered variants with NO host synchronisation, so that a whole ghost refresh costs one host round trip (the
   counts): the bounds land in a 9-double device buffer {lo[3], hi[3], maxExtent[3]} (which the plumbing all-reduces in place
   over NCCL), the selection reads the halo width maxExtent[axis]*(1+1e-9) from that device buffer and leaves
   {nLow, nHigh, cap} (int64) in countsDevice (the table a rank all-gathers).  Lists longer than cap are truncated; the caller
   checks the counts.
   `count` may exceed the number of internal nodes: with reflecting / periodic planes the plane ghosts generated by
   sphb200_reflect_set_ghost_nodes sit right behind the internal nodes and are selected like them (the reference's
   DistributedBoundary comes last in the boundary list and exchanges the ghosts of the other boundaries too); the halo is then
   unpacked behind them (firstGhost = nInternal + nPlaneGhosts) and sphb200_reflect_apply_ghosts / _finalize_derivatives keep
   working on the leading part of the ghost tail. */
int sphb200_node_bounds_device(sphb200_ctx* ctx, size_t count, double* boundsDevice /*[9]*/);
int sphb200_halo_select_device(sphb200_ctx* ctx, int axis, size_t count, double lo, double hi, const double* maxExtentDevice /*[3]*/,
                               uint32_t* sendLowDevice, uint32_t* sendHighDevice, long long* countsDevice /*[3]*/, size_t cap);
/* raw stream handle (cudaStream_t) so the plumbing can order NCCL calls after pack / before unpack */
void* sphb200_stream(sphb200_ctx* ctx);

/* ---- instrumentation ----------------------------------------------------------------------------------------------
   Number of kernels of this library launched on the context since creation, and device time (ms, CUDA events on the
   context's stream) of the last build_pairs / evaluate_derivatives / update_energy calls and of their dominant kernels. */
typedef struct {
  uint64_t launches;
  float ms_build_pairs, ms_evaluate, ms_energy;
  float ms_pair_kernel;        /* the K3 pair-loop kernel alone */
  float ms_neighbor_kernels;   /* K2 count+fill */
  uint64_t directed_edges;     /* sum of neighbour counts (2*internal-internal + internal-ghost pairs) */
  uint32_t stencil_radius;     /* cells a tile walks in each direction at most: 1 when the grid cells are as wide as the largest
                                  kernel extent, > 1 when a heavy tail of extents made the grid follow the typical extent */
  uint32_t fine_walk;          /* 1 if the last build sorted by cells of half the width and walked children (experimental, SPHB200_FINE_WALK=1) */
} sphb200_stats;
int  sphb200_get_stats(sphb200_ctx* ctx, sphb200_stats* out);
/* FP64 FMA throughput microbenchmark on the context's device (roofline denominator; returns TFLOP/s). */
int  sphb200_measure_fp64_peak(sphb200_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
