// cabi_smoke.cpp -- the C ABI of include/sphb200.h called from a plain C++17 translation unit (no Spheral, no Python), the way the
// adapter of INTEGRATION.md section 2 calls it: host fields are AoS std::vector<double> exactly as Field<Dimension, DataType> keeps them
// (internal nodes first, ghosts after), every call in the adapter's order:
//     create -> set_kernel_table (W, gradW, grad2W + the nperh / Wsum lookups) -> set_nodes -> upload_state(position, H) ->
//     build_pairs -> download_pairs / neighbor_counts -> upload_state(the rest) -> evaluate_derivatives -> download_derivs ->
//     download_pair_accelerations -> update_energy_compatible -> download_state(eps) -> copy_DvDx_to_Q -> get_stats -> destroy
// It checks what needs no second implementation (pair-list invariants, antisymmetry of the pair accelerations, sum m DvDt = 0,
// exact total-energy conservation of the compatible update, error reporting) and writes inputs and outputs to a flat binary file;
// tests/test_gpu_cabi_cpp.py feeds the same inputs through the ctypes binding (spheral_b200.engine.Engine) and compares bit for bit.
//   g++ -std=c++17 -O1 -I include tests/cabi_smoke.cpp -o cabi_smoke -L spheral_b200 -lsphb200 -Wl,-rpath,$PWD/spheral_b200
#include "sphb200.h"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static sphb200_ctx* ctx = nullptr;
static void check(int rc, const char* what) {
  if (rc != 0) { std::fprintf(stderr, "FAIL %s: %s\n", what, sphb200_last_error(ctx)); std::exit(2); }
}
static void require(bool ok, const char* what) {
  if (!ok) { std::fprintf(stderr, "FAIL invariant: %s\n", what); std::exit(3); }
}
static void dump(std::FILE* f, const char* name, const void* p, size_t bytes) {
  char tag[32] = {0}; std::strncpy(tag, name, 31);
  const unsigned long long n = bytes;
  std::fwrite(tag, 1, 32, f); std::fwrite(&n, 8, 1, f); std::fwrite(p, 1, bytes, f);
}

int main(int argc, char** argv) {
  const char* outPath = argc > 1 ? argv[1] : "cabi_smoke.bin";
  const int n = argc > 2 ? std::atoi(argv[2]) : 12, ndim = 3;
  const double nPerh = 1.51, dx = 1.0/n;
  const size_t N = (size_t)n*n*n;
  require(sphb200_abi_version() == SPHB200_ABI_VERSION, "ABI version of the library matches the header");

  // ---- a TableKernel payload: TableKernel3d(BSplineKernel3d(), 1000)
  const size_t numPoints = 1000, nc = sphb200_table_ncoef(numPoints);
  std::vector<double> Wc(nc), Gc(nc), G2c(nc), nperhV(2*numPoints), wsumV(2*numPoints);
  double kext = 0, xstep = 0, nperhR[2], wsumR[2]; size_t n1 = 0;
  require(sphb200_table_kernel_build(SPHB200_KERNEL_BSPLINE, ndim, numPoints, 0.25, 64.0, &kext, &xstep, &n1, Wc.data(), Gc.data(), G2c.data(),
                                     nperhV.data(), nperhR, wsumV.data(), wsumR) == 0, "table_kernel_build");
  require(kext == 2.0 && 3*(n1 + 1) == nc, "BSpline table shape");

  // ---- host fields, AoS (Vector 3, SymTensor 6: xx xy xz yy yz zz)
  std::vector<double> pos(3*N), vel(3*N), H(6*N), mass(N), rho(N), eps(N), P(N), cs(N), omega(N);
  unsigned long long seed = 88172645463325252ull;
  auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed >> 11)/9007199254740992.0; };
  for (size_t i = 0; i < N; ++i) {
    const size_t ix = i % n, iy = (i/n) % n, iz = i/((size_t)n*n);
    const double c[3] = {(ix + 0.5)*dx, (iy + 0.5)*dx, (iz + 0.5)*dx};
    for (int a = 0; a < 3; ++a) pos[3*i + a] = c[a] + 0.2*dx*(2.0*rnd() - 1.0);
    vel[3*i] = std::sin(3.0*pos[3*i + 1]); vel[3*i + 1] = std::sin(2.0*pos[3*i + 2]); vel[3*i + 2] = -std::cos(4.0*pos[3*i]);
    const double h = 1.0/(nPerh*dx)*(1.0 + 0.05*(2.0*rnd() - 1.0));
    H[6*i] = h; H[6*i + 1] = 0; H[6*i + 2] = 0; H[6*i + 3] = h; H[6*i + 4] = 0; H[6*i + 5] = h;
    rho[i] = 1.0 + 0.3*std::sin(3.0*pos[3*i])*std::cos(2.0*pos[3*i + 1]);
    mass[i] = rho[i]*dx*dx*dx;
    eps[i] = 1.0 + 0.5*std::cos(2.5*pos[3*i + 1]);
    P[i] = (5.0/3.0 - 1.0)*rho[i]*eps[i];                                   // GammaLawGas (Material/GammaLawGas.cc:185-243)
    cs[i] = std::sqrt((5.0/3.0)*(5.0/3.0 - 1.0)*eps[i]);
    omega[i] = 1.0 + 0.05*(2.0*rnd() - 1.0);
  }

  // ---- the package
  sphb200_options o{};
  o.ndim = ndim; o.compatibleEnergy = 1; o.evolveTotalEnergy = 0; o.XSPH = 1; o.correctVelocityGradient = 1;
  o.epsTensile = 0.0; o.nTensile = 4.0; o.nPerh = nPerh; o.Qkind = SPHB200_Q_MG; o.Cl = 2.0; o.Cq = 2.0; o.eps2 = 1.0e-2;
  o.negligibleSoundSpeed = 1.0e-10; o.etaCritFrac = 1.0; o.etaFoldFrac = 0.2; o.hEvolution = SPHB200_H_SPH; o.hmin = 1.0e-20; o.hmax = 1.0e20;
  o.hydro = SPHB200_HYDRO_SPH;
  sphb200_options bad = o; bad.evolveTotalEnergy = 1;                        // SPHBase.cc: compatible and total energy are exclusive
  sphb200_ctx* none = nullptr;
  require(sphb200_create(&none, 0, &bad) != 0 && std::strstr(sphb200_last_error(nullptr), "cannot simultaneously"), "bad options are rejected with a message");
  check(sphb200_create(&ctx, 0, &o), "create");
  require(sphb200_evaluate_derivatives(ctx, 0.0, 1.0) != 0 && std::strlen(sphb200_last_error(ctx)) > 0, "evaluate before set-up fails loudly");
  check(sphb200_set_kernel_table(ctx, SPHB200_TABLE_W, kext, 0.0, xstep, n1, Wc.data(), Gc.data(), G2c.data(),
                                 numPoints, nperhR[0], nperhR[1], nperhV.data(), numPoints, wsumR[0], wsumR[1], wsumV.data()), "set_kernel_table");
  check(sphb200_set_nodes(ctx, N, 0), "set_nodes");
  sphb200_host_state hs{};
  hs.position = pos.data(); hs.H = H.data();
  check(sphb200_upload_state(ctx, SPHB200_F_POSITION | SPHB200_F_H, &hs), "upload_state(position, H)");
  size_t npairs = 0;
  check(sphb200_build_pairs(ctx, &npairs), "build_pairs");
  require(sphb200_connectivity_valid(ctx) == 1 && npairs > 20*N, "connectivity built");
  std::vector<uint32_t> pi(npairs), pj(npairs), cnt(N);
  check(sphb200_download_pairs(ctx, pi.data(), pj.data(), npairs), "download_pairs");
  check(sphb200_download_neighbor_counts(ctx, cnt.data()), "download_neighbor_counts");
  { std::vector<uint32_t> deg(N, 0);
    for (size_t k = 0; k < npairs; ++k) {
      require(pi[k] < pj[k] && pj[k] < N, "pairs are (i < j)");
      if (k) require(pi[k - 1] < pi[k] || (pi[k - 1] == pi[k] && pj[k - 1] < pj[k]), "pairs sorted by (i, j), no duplicates");
      ++deg[pi[k]]; ++deg[pj[k]];
    }
    for (size_t i = 0; i < N; ++i) require(deg[i] == cnt[i], "neighbour counts = degree in the pair list"); }

  hs = sphb200_host_state{};
  hs.velocity = vel.data(); hs.mass = mass.data(); hs.massDensity = rho.data(); hs.specificThermalEnergy = eps.data();
  hs.pressure = P.data(); hs.soundSpeed = cs.data(); hs.omegaGradh = omega.data();
  check(sphb200_upload_state(ctx, SPHB200_F_VELOCITY | SPHB200_F_MASS | SPHB200_F_RHO | SPHB200_F_EPS | SPHB200_F_PRESSURE |
                                  SPHB200_F_SOUNDSPEED | SPHB200_F_OMEGA, &hs), "upload_state(rest)");
  require(sphb200_connectivity_valid(ctx) == 1, "non-geometric uploads keep the connectivity");
  check(sphb200_evaluate_derivatives(ctx, 0.0, 1.0), "evaluate_derivatives");
  std::vector<double> DxDt(3*N), DrhoDt(N), DvDt(3*N), DepsDt(N), DvDx(9*N), lDvDx(9*N), gradRho(3*N), M(9*N), lM(9*N), rhoSum(N), norm(N),
                      maxQ(N), effQ(N), XW(N), XdV(3*N), DHDt(6*N), Hideal(6*N), m0(N), m1(3*N);
  sphb200_host_derivs hd{};
  hd.DxDt = DxDt.data(); hd.DrhoDt = DrhoDt.data(); hd.DvDt = DvDt.data(); hd.DepsDt = DepsDt.data(); hd.DvDx = DvDx.data();
  hd.localDvDx = lDvDx.data(); hd.gradRho = gradRho.data(); hd.M = M.data(); hd.localM = lM.data(); hd.rhoSum = rhoSum.data();
  hd.normalization = norm.data(); hd.maxViscousPressure = maxQ.data(); hd.effViscousPressure = effQ.data(); hd.XSPHWeightSum = XW.data();
  hd.XSPHDeltaV = XdV.data(); hd.DHDt = DHDt.data(); hd.Hideal = Hideal.data(); hd.massZerothMoment = m0.data(); hd.massFirstMoment = m1.data();
  check(sphb200_download_derivs(ctx, SPHB200_D_ALL, &hd), "download_derivs");
  std::vector<double> pacc(3*npairs);
  check(sphb200_download_pair_accelerations(ctx, pacc.data(), pacc.size()), "download_pair_accelerations");
  // the same evaluation delivered to host fields by the one-call form (chunked pair loop, download overlapped): identical bits
  { std::vector<double> DvDt2(3*N), DvDx2(9*N), DHDt2(6*N), DxDt2(3*N);
    sphb200_host_derivs h2{};
    h2.DvDt = DvDt2.data(); h2.DvDx = DvDx2.data(); h2.DHDt = DHDt2.data(); h2.DxDt = DxDt2.data();
    setenv("SPHB200_E2H_FORCE", "1", 1);                                       // chunk although the problem is small
    check(sphb200_evaluate_derivatives_to_host(ctx, 0.0, 1.0, SPHB200_D_DVDT | SPHB200_D_DVDX | SPHB200_D_DHDT | SPHB200_D_DXDT, &h2), "evaluate_derivatives_to_host");
    unsetenv("SPHB200_E2H_FORCE");
    require(std::memcmp(DvDt2.data(), DvDt.data(), 3*N*sizeof(double)) == 0 && std::memcmp(DvDx2.data(), DvDx.data(), 9*N*sizeof(double)) == 0 &&
            std::memcmp(DHDt2.data(), DHDt.data(), 6*N*sizeof(double)) == 0 && std::memcmp(DxDt2.data(), DxDt.data(), 3*N*sizeof(double)) == 0,
            "evaluate_derivatives_to_host delivers the bits of evaluate_derivatives + download_derivs");
    sphb200_host_derivs none2{};
    require(sphb200_evaluate_derivatives_to_host(ctx, 0.0, 1.0, SPHB200_D_DVDT, &none2) != 0, "a selected field without a destination is an error"); }

  // sum_i m_i DvDt_i = 0 (pairwise antisymmetric forces) and DvDt_i = sum over its pairs of the pair accelerations (SPH.cc:427-430)
  { double mom[3] = {0, 0, 0}, scale = 0.0;
    std::vector<double> acc(3*N, 0.0);
    for (size_t i = 0; i < N; ++i) for (int a = 0; a < 3; ++a) { mom[a] += mass[i]*DvDt[3*i + a]; scale = std::fmax(scale, std::fabs(mass[i]*DvDt[3*i + a])); }
    for (int a = 0; a < 3; ++a) require(std::fabs(mom[a]) <= 1.0e-11*scale*N, "total momentum change vanishes");
    for (size_t k = 0; k < npairs; ++k) for (int a = 0; a < 3; ++a) {
      acc[3*pi[k] + a] += pacc[3*k + a];                                       // pairAccelerations[kk] is the acceleration of i_node ...
      acc[3*pj[k] + a] -= pacc[3*k + a]*mass[pi[k]]/mass[pj[k]];              // ... and j_node gets -m_i/m_j times it
    }
    double worst = 0.0, big = 0.0;
    for (size_t q = 0; q < 3*N; ++q) { worst = std::fmax(worst, std::fabs(acc[q] - DvDt[q])); big = std::fmax(big, std::fabs(DvDt[q])); }
    require(worst <= 1.0e-11*big, "pair accelerations add up to DvDt"); }

  // compatible energy (SpecificThermalEnergyPolicy.cc:47-174): total energy is conserved to round-off for a full step of size dtc
  const double dtc = 1.0e-3;
  check(sphb200_update_energy_compatible(ctx, dtc), "update_energy_compatible");
  std::vector<double> eps1(N);
  { double* fields[14] = {nullptr}; fields[5] = eps1.data();
    check(sphb200_download_state(ctx, SPHB200_F_EPS, fields), "download_state(eps)"); }
  { double dKE = 0.0, dTE = 0.0;
    for (size_t i = 0; i < N; ++i) {
      double k0 = 0, k1 = 0;
      for (int a = 0; a < 3; ++a) { const double v0 = vel[3*i + a], v1 = v0 + dtc*DvDt[3*i + a]; k0 += v0*v0; k1 += v1*v1; }
      dKE += 0.5*mass[i]*(k1 - k0); dTE += mass[i]*(eps1[i] - eps[i]);
    }
    require(std::fabs(dKE + dTE) <= 1.0e-11*std::fmax(std::fabs(dKE), std::fabs(dTE)), "compatible energy update conserves total energy"); }
  check(sphb200_copy_DvDx_to_Q(ctx), "copy_DvDx_to_Q");
  require((sphb200_state_fields_present(ctx) & SPHB200_F_DVDXQ) != 0, "the Q's velocity gradient now lives on the device");
  sphb200_stats st{};
  check(sphb200_get_stats(ctx, &st), "get_stats");
  require(st.launches > 10 && st.directed_edges == 2ull*npairs, "kernels were launched; directed edges = 2 x pairs without ghosts");
  check(sphb200_sync(ctx), "sync");

  std::FILE* f = std::fopen(outPath, "wb");
  require(f != nullptr, "open output file");
  const unsigned long long meta[4] = {(unsigned long long)N, (unsigned long long)npairs, (unsigned long long)ndim, (unsigned long long)n};
  dump(f, "meta", meta, sizeof meta);
#define D(v) dump(f, #v, v.data(), v.size()*sizeof(v[0]))
  D(pos); D(vel); D(H); D(mass); D(rho); D(eps); D(P); D(cs); D(omega);
  D(pi); D(pj); D(cnt); D(DxDt); D(DrhoDt); D(DvDt); D(DepsDt); D(DvDx); D(gradRho); D(M); D(rhoSum); D(norm); D(maxQ); D(effQ); D(XW); D(XdV);
  D(DHDt); D(Hideal); D(m0); D(m1); D(pacc); D(eps1);
#undef D
  std::fclose(f);
  sphb200_destroy(ctx);
  std::printf("cabi_smoke ok: %zu nodes, %zu pairs, %llu kernel launches\n", N, npairs, (unsigned long long)st.launches);
  return 0;
}
