/* ----------------------------------------------------------------------------
 * sph_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, FP64 restatement of the reference (llnl/spheral @ c64796e1)
 * algorithm for the SPH hydro-derivative hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call anything in oracle/.  The product path
 * (spheral_b200/, include/) never links or imports it.
 *
 * The reference itself cannot be compiled in this image (needs Eigen 5, RAJA,
 * CHAI, axom, boost, polytope, silo, MPI -- none vendored), so
 * this restatement is pinned by the reference's own known-answer *properties*
 * (tests/unit/Neighbor/NeighborTestBase.py:192-258 brute-force neighbour sets,
 * tests/unit/SPH/testLinearVelocityGradient.py linear-field DvDx to 5e-5,
 * tests/unit/Kernel/testTableKernel.py:74-90 table vs analytic 1e-3/1e-2,
 * tests/functional/Hydro/Noh/Noh-cylindrical-2d.py:803-808 |dE/E|<1e-13).
 * and by the one numerical golden the reference stores for this path: the L1 / L2
 * / Linf error norms of tests/functional/Hydro/Noh/Noh-planar-1d.py:226-240,
 * which the oracle integrator reproduces to ~1e-6 relative over 1091 steps
 * (tests/test_oracle_noh_planar_1d_golden.py; a 1-D run, hence the D = 1
 * instantiation of the *_dim.inc files).  No per-node golden derivative vectors
 * exist in the reference tree; see DESIGN.md section 7 for what that leaves open.
 *
 * Layout conventions (the reference's Field<DataType> AoS):
 *   Vector     : ndim doubles            (x,y[,z])
 *   SymTensor  : 3-D xx,xy,xz,yy,yz,zz   2-D xx,xy,yy   (GeomSymmetricTensorBase.hh:43-70)
 *   Tensor     : ndim*ndim doubles, row major
 *   nodes      : [0,nInternal) internal, [nInternal,nInternal+nGhost) ghost
 * --------------------------------------------------------------------------*/
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* analytic kernel ids */
enum { ORC_KERNEL_BSPLINE = 0, ORC_KERNEL_WENDLANDC4 = 1, ORC_KERNEL_WENDLANDC2 = 2, ORC_KERNEL_GAUSSIAN = 3,
       ORC_KERNEL_NBSPLINE = 100 /* + order: NBSplineKernel(order), Kernel/NBSplineKernel.cc */ };
/* artificial viscosity ids */
enum { ORC_Q_MG = 0, ORC_Q_LIMITED_MG = 1 };
/* smoothing scale package */
enum { ORC_H_SPH = 0, ORC_H_ASPH = 1, ORC_H_NONE = 2, ORC_H_ASPH_CLASSIC = 3 /* ASPHClassicSmoothingScale (SPHHydros.py:133-134, ASPH = "Classic") */ };

typedef struct {
  int    ndim;                      /* 2 | 3 */
  int    compatibleEnergy;          /* SPHBase.hh:136 */
  int    evolveTotalEnergy;
  int    XSPH;
  int    correctVelocityGradient;
  double epsTensile;                /* SPH.cc:190 */
  double nTensile;                  /* unused by SPH.cc (pow4 hard-wired), kept for API parity */
  double nPerh;                     /* NodeList::nodesPerSmoothingScale */
  /* artificial viscosity (ArtificialViscosityHandle.cc:37-53) */
  int    Qkind;
  double Cl, Cq, eps2, negligibleSoundSpeed;
  int    balsara, linearInExpansion, quadraticInExpansion;
  double etaCritFrac, etaFoldFrac;  /* LimitedMonaghanGingoldViscosity */
  /* smoothing scale */
  int    hEvolution;                /* ORC_H_SPH | ORC_H_ASPH | ORC_H_NONE | ORC_H_ASPH_CLASSIC */
  double hmin, hmax;
  double hminratio;                 /* NodeList::hminratio: ASPHClassicSmoothingScale.cc:254, 296, 355 */
} orc_options;

/* TableKernel payload: three QuadraticInterpolators sharing xmin/xstep/n1
   (Kernel/TableKernel.cc:169-209) + the two CubicHermite nperh lookups. */
typedef struct {
  double kext;
  double xmin, xstep;
  size_t n1;                        /* max interval index; 3*(n1+1) coeffs per table */
  const double* Wcoef;
  const double* gradWcoef;
  const double* grad2Wcoef;
  /* CubicHermite lookups (n values then n gradients), may be NULL/0 if unused */
  size_t nperhN;  double nperhXmin,  nperhXmax,  nperhXstep;  const double* nperhVals;   /* Wsum -> nperh */
  size_t wsumN;   double wsumXmin,   wsumXmax,   wsumXstep;   const double* wsumVals;    /* nperh -> Wsum */
} orc_table;

typedef struct {                    /* inputs, AoS, length nInternal+nGhost */
  const double *pos, *vel, *H, *mass, *rho, *P, *cs, *omega;
  const double *DvDxQ;              /* optional (LimitedMG / Balsara), Tensor */
  const double *fCl, *fCq;          /* optional multipliers */
} orc_state;

typedef struct {                    /* outputs, AoS, length nInternal+nGhost (ghost entries left 0) */
  double *DxDt, *DrhoDt, *DvDt, *DepsDt, *DvDx, *localDvDx, *gradRho, *M, *localM;
  double *rhoSum, *normalization, *maxViscousPressure, *effViscousPressure;
  double *XSPHWeightSum, *XSPHDeltaV;
  double *DHDt, *Hideal, *massZerothMoment, *massFirstMoment;
  double *pairAccelerations;        /* npairs * ndim, may be NULL unless compatibleEnergy */
} orc_derivs;

/* ---- kernels and tables --------------------------------------------------*/
double orc_kernel_extent(int kind, int ndim);
void   orc_kernel_analytic(int kind, int ndim, double eta, double* W, double* gradW, double* grad2W);
/* QuadraticInterpolator::initialize (Utilities/QuadraticInterpolator.cc:59-100) */
void   orc_quadratic_fit(double xmin, double xmax, size_t n, const double* yvals, double* coeffs, size_t* n1, double* xstep);
/* table sizes: returns number of coefficients per table ( 3*(n1+1) ) for numPoints */
size_t orc_table_ncoef(size_t numPoints);
int    orc_table_build(int kind, int ndim, size_t numPoints, double* Wc, double* gWc, double* g2Wc,
                       double* kext, size_t* n1, double* xstep);
/* nperh lookups (TableKernel.cc:196-208): vals arrays each 2*numPoints doubles */
int    orc_table_build_nperh(const orc_table* t, int ndim, size_t numPoints, double minNperh, double maxNperh,
                             double* wsumVals, double* wsumRange /*[xmin,xmax]*/,
                             double* nperhVals, double* nperhRange);
void   orc_table_eval(const orc_table* t, double eta, double Hdet, double* W, double* gW);
double orc_cubic_hermite_eval(size_t n, double xmin, double xmax, double xstep, const double* vals, double x);

/* ---- neighbour pairs (ConnectivityMap.cc:912-931, 1023-1029) ---------------*/
/* returns npairs; writes up to cap sorted (i<j) pairs; counts[i] = numNeighborsForNode (internal i) */
size_t orc_pairs_bruteforce(int ndim, size_t nInt, size_t nGhost, const double* pos, const double* H,
                            double kext, uint32_t* pi, uint32_t* pj, size_t cap, uint32_t* counts);
size_t orc_pairs_cells(int ndim, size_t nInt, size_t nGhost, const double* pos, const double* H,
                       double kext, uint32_t* pi, uint32_t* pj, size_t cap, uint32_t* counts);

/* ---- derivatives ------------------------------------------------------------*/
/* SPH<Dim>::evaluateDerivativesImpl (SPH/SPH.cc:165-555) followed by the
   smoothing-scale sub-package (SPHSmoothingScale.cc:101-275 | ASPHSmoothingScale.cc:110-147).
   nthreads<=1: serial, pair order; else OpenMP with per-thread scratch copies
   (the reference's threadCopy/threadReduce strategy, FieldListInline.hh:1394-1443). */
int orc_evaluate_derivatives(const orc_options* o, const orc_table* W, const orc_table* WQ,
                             size_t nInt, size_t nGhost, const orc_state* s,
                             size_t npairs, const uint32_t* pi, const uint32_t* pj,
                             const uint32_t* numNeighbors, orc_derivs* d, int nthreads);

/* SpecificThermalEnergyPolicy::update (Hydro/SpecificThermalEnergyPolicy.cc:47-174) : eps += ... */
int orc_update_energy_compatible(int ndim, size_t nInt, size_t nGhost, const double* mass, const double* vel,
                                 const double* DvDt, const double* DepsDt0, size_t npairs,
                                 const uint32_t* pi, const uint32_t* pj, const double* pairAccelerations,
                                 double multiplier, double* eps);

/* ---- CRKSPH (BASELINE config 4): LinearOrder reproducing kernels, RKSumVolume ---------------------
   corrections: (1+ndim)*(1+ndim) doubles per node = C[1+ndim] then dC_d[1+ndim] per direction
   (RKCoefficients layout, RK/RKUtilitiesInline.hh:58-105).  vol / corr / rho are written for internal
   nodes only; ghost entries must be supplied by the caller (the reference's boundary conditions). */
int orc_crk_sum_volume(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* H,
                       size_t npairs, const uint32_t* pi, const uint32_t* pj, double* vol);      /* RK/computeRKSumVolume.cc:33-116 */
int orc_crk_corrections(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* H,
                        const double* vol, size_t npairs, const uint32_t* pi, const uint32_t* pj, double* corr);  /* RK/RKUtilities.cc:252-491 */
int orc_crk_sum_density(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* mass,
                        const double* vol, const double* H, size_t npairs, const uint32_t* pi, const uint32_t* pj,
                        double rhoMin, double rhoMax, double* rho);                              /* CRKSPH/computeCRKSPHSumMassDensity.cc:21-131 */
int orc_crk_evaluate_derivatives(const orc_options* o, const orc_table* W, size_t nInt, size_t nGhost,
                                 const orc_state* s, const double* vol, const double* corr,
                                 size_t npairs, const uint32_t* pi, const uint32_t* pj, orc_derivs* d);  /* CRKSPH/CRKSPH.cc:176-440 */
void orc_rk_kernel_grad(int ndim, const orc_table* W, const double* x, const double* H, const double* corr,
                        double* WR, double* gradWR);                                             /* RK/RKUtilities.cc:180-209 */

/* ---- per-step callers of the derivative path (step_oracle_dim.inc) -----------------------
   orc_step_options: what the reference keeps on the FluidNodeList / EquationOfState / Integrator / SmoothingScaleBase. */
typedef struct {
  double gamma;                              /* GammaLawGas (Material/GammaLawGas.cc) */
  double minimumPressure, maximumPressure, externalPressure;
  int    minPressureType;                    /* 0 PressureFloor | 1 ZeroPressure (EquationOfStateInline.hh:98-106) */
  double rhoMin, rhoMax;                     /* FluidNodeList::rhoMin/rhoMax */
  double hminratio;                          /* NodeList::hminratio (IncrementASPHHtensor.cc:58) */
  int    HEvolution;                         /* 0 IdealH | 1 IntegrateH | 2 FixedH (SmoothingScaleBase.hh) */
  double cfl;                                /* GenericHydro::cfl */
  int    useVelocityMagnitudeForDt;
} orc_step_options;
int orc_sum_mass_density(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* mass,
                         const double* H, size_t npairs, const uint32_t* pi, const uint32_t* pj, double* rho);
int orc_omega_gradh(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* H,
                    size_t npairs, const uint32_t* pi, const uint32_t* pj, const uint32_t* numNeighbors, double* omega);
void orc_eos_gamma_law(const orc_step_options* so, size_t n, const double* rho, const double* eps, double* P, double* cs);
int orc_state_update(const orc_options* o, const orc_step_options* so, size_t nInt, size_t nGhost, double multiplier,
                     int timeAdvanceOnly, int epsDone, const orc_derivs* d,
                     double* pos, double* vel, double* H, double* rho, double* eps, double* P, double* cs);
double orc_hydro_dt(const orc_options* o, const orc_step_options* so, size_t nInt, const double* vel, const double* H,
                    const double* rho, const double* cs, const orc_derivs* d, size_t npairs, const uint32_t* pi,
                    const uint32_t* pj, int* reason, uint32_t* node);
void orc_sym_bound(int ndim, double* H, double minv, double maxv);   /* test hook: eigenvalue clamp of one SymTensor */

#ifdef __cplusplus
}
#endif
#endif
