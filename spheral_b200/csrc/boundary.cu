// boundary.cu -- reflecting-plane boundary conditions on the device: ghost-node generation,
// ghost-value refresh and enforcement, so that the stock Noh / Sedov set-ups (reflecting planes through the origin) run with the
// state resident on the GPU.
//
//   sphb200_reflect_set_ghost_nodes   ReflectingBoundary / PlanarBoundary::setGhostNodes, one plane after the other so that later
//                                     planes mirror the ghosts of earlier ones (Integrator/Integrator.cc:415-424);
//                                     control nodes by findNodesTouchingThroughPlanes (Boundary/findNodesTouchingThroughPlanes.cc),
//                                     ghost positions by mapPositionThroughPlanes (Boundary/mapPositionThroughPlanes.hh:17-27)
//   sphb200_reflect_apply_ghosts      ReflectingBoundary::applyGhostBoundary (Boundary/ReflectingBoundary.cc:182-250): scalars
//                                     copied, vectors R v, tensors R (T R), symmetric tensors (R (H R)).Symmetric(),
//                                     R = I - 2 n (x) n (Utilities/planarReflectingOperator.hh:14-19)
//   sphb200_reflect_enforce           PlanarBoundary::setViolationNodes / enforceBoundary (Boundary/PlanarBoundary.cc:153-190,
//                                     ReflectingBoundary.cc:255-330): an internal node that crossed a plane is mapped back and its
//                                     velocity reflected
//
// Control lists are built in ascending node order (flags + scan + scatter), so ghost indices are a fixed function of the input.
#include "sphb200_internal.cuh"
#include "sym_eigen.cuh"
#include <algorithm>
#include <cmath>

namespace {

constexpr int RB = 256;

// One planar boundary (PlanarBoundary.hh): enter plane {p, n} and exit plane {px, nx}, unit normals pointing into the domain.
// Reflecting: exit == enter.  Periodic: a PeriodicBoundary is two of these, (plane1 -> plane2) and (plane2 -> plane1)
// (PeriodicBoundary.cc:60-62).  Control nodes sit within a kernel extent of the EXIT plane; their ghosts appear behind the ENTER
// plane at mapPosition(r, exit, enter) = closestPointOnPlane_enter(r) - signedDistance_exit(r) n_enter (PlanarBoundary.cc:318-320,
// mapPositionThroughPlanes.hh:17-27).
struct Plane { double p[3]; double n[3]; double px[3]; double nx[3]; int periodic; };

template <int DIM> __device__ __forceinline__ double signed_distance(const Plane& pl, const double* r) {      // to the enter plane
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; ++q) s += (r[q] - pl.p[q])*pl.n[q];
  return s;
}
template <int DIM> __device__ __forceinline__ double signed_distance_exit(const Plane& pl, const double* r) {
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; ++q) s += (r[q] - pl.px[q])*pl.nx[q];
  return s;
}

// pass 1: hmax = largest 1/min-eigenvalue(H_i) among nodes closer than kext*hmax_i to either plane
// (findNodesTouchingThroughPlanes.cc, active branch)
template <int DIM>
__global__ void __launch_bounds__(RB) k_reflect_hmax(const double* __restrict__ pos, const double* __restrict__ H, size_t n, Plane pl,
                                                     double kext, unsigned long long* __restrict__ hmaxBits) {
  constexpr int NS = Dm<DIM>::NS;
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  double v = 0.0;
  if (i < n) {
    double Hi[NS], r[DIM];
#pragma unroll
    for (int q = 0; q < NS; ++q) Hi[q] = H[i*NS + q];
#pragma unroll
    for (int q = 0; q < DIM; ++q) r[q] = pos[i*DIM + q];
    const double dmin = fmin(fabs(signed_distance<DIM>(pl, r)), fabs(signed_distance_exit<DIM>(pl, r)));
    // Gershgorin: lambda_min >= min over rows of (diagonal - sum |off-diagonal|).  Where that bound is positive it caps h_i from
    // above, and a node farther from the plane than kext times the cap cannot pass the test below: the eigenvalue (the expensive
    // part of this pass: 0.17 ms per plane at 8 M nodes) is only computed for the few layers of nodes next to the plane.
    double glo;
    if (DIM == 3) glo = fmin(Hi[0] - fabs(Hi[1]) - fabs(Hi[2]), fmin(Hi[3] - fabs(Hi[1]) - fabs(Hi[4]), Hi[5] - fabs(Hi[2]) - fabs(Hi[4])));
    else glo = fmin(Hi[0] - fabs(Hi[1]), Hi[2] - fabs(Hi[1]));
    if (!(glo > 0.0) || dmin*glo < kext*(1.0 + 1.0e-12)) {
      const double hi = 1.0/sym_min_eigenvalue<DIM>(Hi);
      if (dmin < kext*hi) v = hi;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax(hmaxBits, (unsigned long long)__double_as_longlong(v));   // positive doubles order as integers
}
// pass 2: control flags 0 <= signedDistance_exit/hmax <= kext
template <int DIM>
__global__ void __launch_bounds__(RB) k_reflect_flags(const double* __restrict__ pos, size_t n, Plane pl, double kext,
                                                      const unsigned long long* __restrict__ hmaxBits, uint32_t* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  const double hmax = __longlong_as_double((long long)*hmaxBits);
  double r[DIM];
#pragma unroll
  for (int q = 0; q < DIM; ++q) r[q] = pos[i*DIM + q];
  const double x = signed_distance_exit<DIM>(pl, r)/hmax;
  flags[i] = (hmax > 0.0 && x >= 0.0 && x <= kext) ? 1u : 0u;
}
__global__ void __launch_bounds__(RB) k_reflect_scatter(const uint32_t* __restrict__ scan, size_t n, uint32_t* __restrict__ ctl, size_t cap) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  if (scan[i + 1] != scan[i] && scan[i] < cap) ctl[scan[i]] = (uint32_t)i;
}

// one field of the ghosts [first, first+count) of a plane from their control nodes
//   kind 0 scalar copy | 1 position (mirror) | 2 vector R v | 3 tensor R (T R) | 4 symmetric tensor (R (H R)).Symmetric()
//   kind 5 RK coefficients of linear order {A, B_k | dA/dx_d, dB_k/dx_d}: ReflectingBoundary::applyGhostBoundary(Field<RKCoefficients>)
//          (Boundary/ReflectingBoundary.cc:403-432) applies RKUtilities::getTransformationMatrix(R) (RK/RKUtilities.cc:637-715):
//          A' = A, B' = R.B, (grad A)' = R.grad A, (grad B)' = R.(grad B).R; periodic boundaries copy
// all selected fields of a plane's ghosts in ONE launch (blockIdx.y = field): a launch per field made plane-ghost generation
// launch-bound on small ranks (9 fields x 3 planes per step)
struct FillSet { double* f[S_COUNT]; int kind[S_COUNT]; int width[S_COUNT]; int n; };
template <int DIM>
__device__ __forceinline__ void reflect_fill_one(double* __restrict__ f, int kind, int width, const uint32_t* __restrict__ ctl,
                                                 size_t first, size_t count, const Plane& pl) {
  const size_t k = (size_t)blockIdx.x*RB + threadIdx.x;
  if (k >= count) return;
  const size_t c = ctl[k], g = first + k;
  const double* s = f + c*(size_t)width;
  double* d = f + g*(size_t)width;
  if (kind == 0 || (pl.periodic && kind != 1)) { for (int q = 0; q < width; ++q) d[q] = s[q]; return; }
  double R[DIM][DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a)
#pragma unroll
    for (int b = 0; b < DIM; ++b) R[a][b] = (a == b ? 1.0 : 0.0) - 2.0*pl.n[a]*pl.n[b];
  if (kind == 5) {
    constexpr int PS = DIM + 1;                      // polynomial size of linear order; block 0: coefficients, block 1+e: d/dx_e
    double in[PS*PS];
#pragma unroll
    for (int q = 0; q < PS*PS; ++q) in[q] = s[q];    // the ghost may alias nothing of its control, but read first all the same
    d[0] = in[0];
#pragma unroll
    for (int a = 0; a < DIM; ++a) { double t = 0.0, u = 0.0;
#pragma unroll
      for (int b = 0; b < DIM; ++b) { t += R[a][b]*in[1 + b]; u += R[a][b]*in[PS*(1 + b)]; }
      d[1 + a] = t;                                  // B' = R.B
      d[PS*(1 + a)] = u; }                           // (grad A)' = R.grad A
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int k = 0; k < DIM; ++k) { double t = 0.0;
#pragma unroll
        for (int e = 0; e < DIM; ++e)
#pragma unroll
          for (int m = 0; m < DIM; ++m) t += R[a][e]*in[PS*(1 + e) + 1 + m]*R[m][k];
        d[PS*(1 + a) + 1 + k] = t; }                 // (grad B)' = R.(grad B).R
    return;
  }
  if (kind == 1) {                                   // closestPointOnPlane_enter(r) - signedDistance_exit(r) n_enter
    double r[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) r[q] = s[q];
    const double sde = signed_distance<DIM>(pl, r), sdx = signed_distance_exit<DIM>(pl, r);
#pragma unroll
    for (int q = 0; q < DIM; ++q) d[q] = (r[q] - sde*pl.n[q]) - sdx*pl.n[q];
  } else if (kind == 2) {
#pragma unroll
    for (int a = 0; a < DIM; ++a) { double t = 0.0;
#pragma unroll
      for (int b = 0; b < DIM; ++b) t += R[a][b]*s[b];
      d[a] = t; }
  } else {
    double T[DIM][DIM], TR[DIM][DIM], O[DIM][DIM];
    if (kind == 3) {
#pragma unroll
      for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = 0; b < DIM; ++b) T[a][b] = s[a*DIM + b];
    } else if (DIM == 3) {
      T[0][0] = s[0]; T[0][1] = T[1][0] = s[1]; T[0][2] = T[2][0] = s[2]; T[1][1] = s[3]; T[1][2] = T[2][1] = s[4]; T[2][2] = s[5];
    } else {
      T[0][0] = s[0]; T[0][1] = T[1][0] = s[1]; T[1][1] = s[2];
    }
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) { double t = 0.0;
#pragma unroll
        for (int e = 0; e < DIM; ++e) t += T[a][e]*R[e][b];
        TR[a][b] = t; }
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) { double t = 0.0;
#pragma unroll
        for (int e = 0; e < DIM; ++e) t += R[a][e]*TR[e][b];
        O[a][b] = t; }
    if (kind == 3) {
#pragma unroll
      for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = 0; b < DIM; ++b) d[a*DIM + b] = O[a][b];
    } else if (DIM == 3) {
      d[0] = O[0][0]; d[1] = 0.5*(O[0][1] + O[1][0]); d[2] = 0.5*(O[0][2] + O[2][0]); d[3] = O[1][1]; d[4] = 0.5*(O[1][2] + O[2][1]); d[5] = O[2][2];
    } else {
      d[0] = O[0][0]; d[1] = 0.5*(O[0][1] + O[1][0]); d[2] = O[1][1];
    }
  }
}
template <int DIM>
__global__ void __launch_bounds__(RB) k_reflect_fill(FillSet fs, const uint32_t* __restrict__ ctl, size_t first, size_t count, Plane pl) {
  const int q = blockIdx.y;
  reflect_fill_one<DIM>(fs.f[q], fs.kind[q], fs.width[q], ctl, first, count, pl);
}

// enforceBoundary: internal nodes behind a plane (signed distance < 0) are mirrored back, their velocity reflected and -- for a
// reflecting plane -- their H tensor as well: PlanarBoundary::updateViolationNodes (PlanarBoundary.cc:179-195) calls
// enforceBoundary on the position, the velocity AND the H field; ReflectingBoundary::enforceBoundary(Field<SymTensor>)
// (ReflectingBoundary.cc:493-501) maps H -> (R H R).Symmetric(), which only matters for anisotropic (ASPH) tensors.
template <int DIM>
__global__ void __launch_bounds__(RB) k_reflect_enforce(double* __restrict__ pos, double* __restrict__ vel, double* __restrict__ H, size_t nInt, Plane pl,
                                                        unsigned long long* __restrict__ nViolations) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= nInt) return;
  double r[DIM];
#pragma unroll
  for (int q = 0; q < DIM; ++q) r[q] = pos[i*DIM + q];
  const double sd = signed_distance<DIM>(pl, r);
  if (sd >= 0.0) return;                             // not below the enter plane (PlanarBoundary.cc:165-167)
  // mapPosition(r, enter, exit) = closestPointOnPlane_exit(r) - signedDistance_enter(r) n_exit (PlanarBoundary.cc:186)
  const double sdx = signed_distance_exit<DIM>(pl, r);
  double vn = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; ++q) { pos[i*DIM + q] = (r[q] - sdx*pl.nx[q]) - sd*pl.nx[q]; vn += vel[i*DIM + q]*pl.n[q]; }
  if (!pl.periodic) {                                // ReflectingBoundary::enforceBoundary(velocity): v -> R v
#pragma unroll
    for (int q = 0; q < DIM; ++q) vel[i*DIM + q] -= 2.0*vn*pl.n[q];
    if (H) {                                         // H -> (R H R).Symmetric(), R = I - 2 n (x) n
      constexpr int NS = Dm<DIM>::NS;
      double T[DIM][DIM], O[DIM][DIM], Hn[DIM], nHn = 0.0;
      const double* s = H + i*NS;
      if (DIM == 3) { T[0][0] = s[0]; T[0][1] = T[1][0] = s[1]; T[0][DIM - 1] = T[DIM - 1][0] = s[2]; T[1][1] = s[3]; T[1][DIM - 1] = T[DIM - 1][1] = s[4]; T[DIM - 1][DIM - 1] = s[5]; }
      else { T[0][0] = s[0]; T[0][1] = T[1][0] = s[1]; T[1][1] = s[2]; }
#pragma unroll
      for (int a = 0; a < DIM; ++a) { double t = 0.0;
#pragma unroll
        for (int b = 0; b < DIM; ++b) t += T[a][b]*pl.n[b];
        Hn[a] = t; nHn += pl.n[a]*t; }
      // R H R = H - 2 n (Hn)^T - 2 (Hn) n^T + 4 (n.Hn) n n^T   (H symmetric: the result is symmetric up to round-off)
#pragma unroll
      for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = 0; b < DIM; ++b) O[a][b] = T[a][b] - 2.0*pl.n[a]*Hn[b] - 2.0*Hn[a]*pl.n[b] + 4.0*nHn*pl.n[a]*pl.n[b];
      double* d = H + i*NS;
      if (DIM == 3) { d[0] = O[0][0]; d[1] = 0.5*(O[0][1] + O[1][0]); d[2] = 0.5*(O[0][DIM - 1] + O[DIM - 1][0]); d[3] = O[1][1]; d[4] = 0.5*(O[1][DIM - 1] + O[DIM - 1][1]); d[5] = O[DIM - 1][DIM - 1]; }
      else { d[0] = O[0][0]; d[1] = 0.5*(O[0][1] + O[1][0]); d[2] = O[1][1]; }
    }
  }
  atomicAdd(nViolations, 1ull);
}

// finalizeDerivatives: ghost values of the acceleration (R a) and of DepsDt (copy) in the sorted, component-major derivative arrays
template <int DIM>
__global__ void __launch_bounds__(RB) k_reflect_derivs(double* __restrict__ DvDt, double* __restrict__ DepsDt, size_t cap,
                                                       const uint32_t* __restrict__ inv, const uint32_t* __restrict__ ctl,
                                                       size_t first, size_t count, Plane pl) {
  const size_t k = (size_t)blockIdx.x*RB + threadIdx.x;
  if (k >= count) return;
  const size_t sc = inv[ctl[k]], sg = inv[first + k];
  double a[DIM], an = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; ++q) { a[q] = DvDt[(size_t)q*cap + sc]; an += a[q]*pl.n[q]; }
  if (pl.periodic) an = 0.0;                         // periodic images carry the control node's acceleration unchanged
#pragma unroll
  for (int q = 0; q < DIM; ++q) DvDt[(size_t)q*cap + sg] = a[q] - 2.0*an*pl.n[q];
  DepsDt[sg] = DepsDt[sc];
}

int field_kind(int ndim, int slot) {
  switch (slot) {
    case S_POS: return 1; case S_VEL: return 2; case S_H: return 4; case S_DVDXQ: return 3; case S_RKCORR: return 5;
    default: return 0;
  }
}
Plane plane_of(const sphb200_ctx* c, int p) {
  Plane pl{};
  for (int q = 0; q < 3; ++q) { pl.p[q] = c->planes[12*p + q]; pl.n[q] = c->planes[12*p + 3 + q]; pl.px[q] = c->planes[12*p + 6 + q]; pl.nx[q] = c->planes[12*p + 9 + q]; }
  pl.periodic = c->planeKind[p];
  return pl;
}
int fill_plane(sphb200_ctx* c, int p, unsigned mask) {
  const size_t first = c->planeFirst[p], count = c->planeCount[p];
  if (count == 0) return 0;
  const Plane pl = plane_of(c, p);
  const unsigned nb = (unsigned)((count + RB - 1)/RB);
  FillSet fs{};
  for (int s = 0; s < S_COUNT; ++s) {
    if (!(mask & (1u << s)) || !c->have[s] || !c->api[s]) continue;
    fs.f[fs.n] = c->api[s]; fs.kind[fs.n] = field_kind(c->ndim, s); fs.width[fs.n] = sphb200_state_width(c->ndim, s); ++fs.n;
  }
  if (fs.n == 0) return 0;
  const dim3 grid(nb, (unsigned)fs.n);
  if (c->ndim == 3) k_reflect_fill<3><<<grid, RB, 0, c->stream>>>(fs, c->planeCtl[p], first, count, pl);
  else              k_reflect_fill<2><<<grid, RB, 0, c->stream>>>(fs, c->planeCtl[p], first, count, pl);
  KERNEL_CHECK(c, "k_reflect_fill");
  return 0;
}

}  // namespace

extern "C" {

int sphb200_boundary_configure(sphb200_ctx* c, int nBoundaries, const int* kinds, const double* enterPoints, const double* enterNormals,
                               const double* exitPoints, const double* exitNormals) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (nBoundaries < 0 || nBoundaries > SPHB200_MAX_PLANES) return sphb200_fail(c, "boundary_configure: between 0 and 6 planar boundaries are supported");
  if (nBoundaries && (!kinds || !enterPoints || !enterNormals)) return sphb200_fail(c, "boundary_configure: null boundary data");
  for (int p = 0; p < nBoundaries; ++p) {
    const bool periodic = kinds[p] == SPHB200_BOUNDARY_PERIODIC;
    if (kinds[p] != SPHB200_BOUNDARY_REFLECTING && !periodic) return sphb200_fail(c, "boundary_configure: unknown boundary kind");
    if (periodic && (!exitPoints || !exitNormals)) return sphb200_fail(c, "boundary_configure: a periodic boundary needs its exit plane");
    const double* pts[2] = {enterPoints, periodic ? exitPoints : enterPoints};
    const double* nrm[2] = {enterNormals, periodic ? exitNormals : enterNormals};
    for (int side = 0; side < 2; ++side) {
      double nn = 0.0;
      for (int q = 0; q < c->ndim; ++q) nn += nrm[side][p*c->ndim + q]*nrm[side][p*c->ndim + q];
      if (!(nn > 0.0)) return sphb200_fail(c, "boundary_configure: zero plane normal");
      nn = std::sqrt(nn);
      for (int q = 0; q < 3; ++q) {
        c->planes[12*p + 6*side + q] = q < c->ndim ? pts[side][p*c->ndim + q] : 0.0;
        c->planes[12*p + 6*side + 3 + q] = q < c->ndim ? nrm[side][p*c->ndim + q]/nn : 0.0;      // unit normal, pointing into the domain
      }
    }
    c->planeKind[p] = periodic ? 1 : 0;
    c->planeFirst[p] = c->planeCount[p] = 0;
  }
  c->nPlanes = nBoundaries;
  return 0;
}

int sphb200_reflect_configure(sphb200_ctx* c, int nPlanes, const double* points, const double* normals) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (nPlanes < 0 || nPlanes > SPHB200_MAX_PLANES) return sphb200_fail(c, "reflect_configure: between 0 and 6 planes are supported");
  int kinds[SPHB200_MAX_PLANES] = {0};
  return sphb200_boundary_configure(c, nPlanes, kinds, points, normals, nullptr, nullptr);
}

int sphb200_reflect_set_ghost_nodes(sphb200_ctx* c, size_t* nGhostOut) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  // Which nodes become control nodes, and where their ghosts go, depends on positions and H alone: only those uploads are waited for.
  // If other fields are still in flight on the copy stream the ghosts get whatever values are on the device now and are refilled when
  // the first call that needs every field joins the uploads (sphb200_join_uploads) -- the neighbour build in between needs geometry only.
  const bool restInFlight = c->pendRestUp;
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, false)) return 1;
  if (!c->have[S_POS] || !c->have[S_H]) return sphb200_fail(c, "reflect_set_ghost_nodes: position and H must be on the device");
  if (!c->W.set) return sphb200_fail(c, "reflect_set_ghost_nodes: kernel table not set (need the kernel extent)");
  const double kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  // start from the internal nodes alone: previous ghosts are discarded (Integrator::setGhostNodes, Integrator.cc:389-397)
  if (sphb200_set_nodes(c, c->nInt, 0)) return 1;
  unsigned mask = 0;
  for (int s = 0; s < S_COUNT; ++s) if (c->have[s] && c->api[s]) mask |= 1u << s;
  for (int p = 0; p < c->nPlanes; ++p) {
    const size_t n = c->n;
    const Plane pl = plane_of(c, p);
    c->planeFirst[p] = n; c->planeCount[p] = 0;
    if (n == 0) continue;
    if (sphb200_ensure(c, c->planeCtl[p], c->planeCtlCap[p], n + n/8 + 32)) return 1;
    // flags live in the download staging area: n+1 counters + one 64-bit cell for hmax
    const size_t need = (n + 4)*sizeof(uint32_t) + 16;
    if (need > c->stageBytes) {
      CU_CHECK(c, cudaStreamSynchronize(c->stream));
      if (c->stage) cudaFree(c->stage);
      c->stage = nullptr; c->stageBytes = 0;
      CU_CHECK(c, cudaMalloc((void**)&c->stage, need + need/8));
      c->stageBytes = need + need/8;
    }
    unsigned long long* hmaxBits = (unsigned long long*)c->stage;
    uint32_t* flags = (uint32_t*)(hmaxBits + 2);
    CU_CHECK(c, cudaMemsetAsync(hmaxBits, 0, 16, c->stream));
    const unsigned nb = (unsigned)((n + RB - 1)/RB);
    if (c->ndim == 3) k_reflect_hmax<3><<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], n, pl, kext, hmaxBits);
    else              k_reflect_hmax<2><<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], n, pl, kext, hmaxBits);
    KERNEL_CHECK(c, "k_reflect_hmax");
    if (c->ndim == 3) k_reflect_flags<3><<<nb, RB, 0, c->stream>>>(c->api[S_POS], n, pl, kext, hmaxBits, flags);
    else              k_reflect_flags<2><<<nb, RB, 0, c->stream>>>(c->api[S_POS], n, pl, kext, hmaxBits, flags);
    KERNEL_CHECK(c, "k_reflect_flags");
    if (sphb200_scan_u32(c, flags, flags, n)) return 1;
    k_reflect_scatter<<<nb, RB, 0, c->stream>>>(flags, n, c->planeCtl[p], c->planeCtlCap[p]);
    KERNEL_CHECK(c, "k_reflect_scatter");
    uint32_t count = 0;
    CU_CHECK(c, cudaMemcpyAsync(&count, flags + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    if (count > c->planeCtlCap[p]) return sphb200_fail(c, "reflect_set_ghost_nodes: internal control list too small");
    c->planeCount[p] = count;
    if (count == 0) continue;
    if (sphb200_set_nodes(c, c->nInt, c->nGhost + count)) return 1;       // state of the existing nodes is kept
    if (fill_plane(c, p, mask)) return 1;
  }
  c->sortValid = c->rowsValid = c->pairsValid = false;
  c->ghostRefillPending = restInFlight && c->nGhost > 0;
  if (nGhostOut) *nGhostOut = c->nGhost;
  return 0;
}

int sphb200_reflect_apply_ghosts(sphb200_ctx* c, unsigned fieldMask) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  size_t total = 0;
  for (int p = 0; p < c->nPlanes; ++p) total += c->planeCount[p];
  // the plane ghosts lead the ghost tail; ghosts of a later boundary (the slab halo of a decomposed run, which also carries copies
  // of the neighbours' plane ghosts -- DistributedBoundary comes last in the reference's boundary list) may follow them
  if (total > c->nGhost) return sphb200_fail(c, "reflect_apply_ghosts: the ghost nodes were not generated by reflect_set_ghost_nodes (or the node count changed since)");
  for (int p = 0; p < c->nPlanes; ++p) if (fill_plane(c, p, fieldMask)) return 1;      // in order: later planes mirror earlier ghosts
  // ghost values changed under a fixed connectivity (the reference refreshes ghosts mid-step without a neighbour update)
  c->rowsValid = false;
  return 0;
}

int sphb200_reflect_finalize_derivatives(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->derivsValid || !c->pairsValid) return sphb200_fail(c, "reflect_finalize_derivatives: derivatives have not been evaluated on the current connectivity");
  size_t total = 0;
  for (int p = 0; p < c->nPlanes; ++p) total += c->planeCount[p];
  if (total > c->nGhost) return sphb200_fail(c, "reflect_finalize_derivatives: the ghost nodes were not generated by reflect_set_ghost_nodes");
  if (total == 0) return 0;
  if (sphb200_inverse_perm(c)) return 1;
  for (int p = 0; p < c->nPlanes; ++p) {
    const size_t count = c->planeCount[p];
    if (count == 0) continue;
    const Plane pl = plane_of(c, p);
    const unsigned nb = (unsigned)((count + RB - 1)/RB);
    if (c->ndim == 3) k_reflect_derivs<3><<<nb, RB, 0, c->stream>>>(c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], c->cap, c->invPerm, c->planeCtl[p], c->planeFirst[p], count, pl);
    else              k_reflect_derivs<2><<<nb, RB, 0, c->stream>>>(c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], c->cap, c->invPerm, c->planeCtl[p], c->planeFirst[p], count, pl);
    KERNEL_CHECK(c, "k_reflect_derivs");
  }
  return 0;
}

int sphb200_reflect_enforce(sphb200_ctx* c, size_t* nViolations) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->have[S_POS] || !c->have[S_VEL]) return sphb200_fail(c, "reflect_enforce: position and velocity must be on the device");
  if (nViolations) *nViolations = 0;
  if (c->nInt == 0 || c->nPlanes == 0) return 0;
  CU_CHECK(c, cudaMemsetAsync(c->counters + 7, 0, sizeof(unsigned long long), c->stream));
  const unsigned nb = (unsigned)((c->nInt + RB - 1)/RB);
  for (int p = 0; p < c->nPlanes; ++p) {
    const Plane pl = plane_of(c, p);
    if (c->ndim == 3) k_reflect_enforce<3><<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_VEL], c->have[S_H] ? c->api[S_H] : nullptr, c->nInt, pl, c->counters + 7);
    else              k_reflect_enforce<2><<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_VEL], c->have[S_H] ? c->api[S_H] : nullptr, c->nInt, pl, c->counters + 7);
    KERNEL_CHECK(c, "k_reflect_enforce");
  }
  c->rowsValid = false;
  if (nViolations) {
    unsigned long long v = 0;
    CU_CHECK(c, cudaMemcpyAsync(&v, c->counters + 7, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    *nViolations = (size_t)v;
    if (v) { c->sortValid = false; c->pairsValid = false; }
  }
  return 0;
}

}  // extern "C"
