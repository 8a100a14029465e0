"""The reference's own integrated regression test for this path, run with the ORACLE integrator (CPU, test infrastructure):
tests/functional/Hydro/Noh/Noh-planar-1d.py -- the planar Noh problem in 1-D with the stock SPH hydro.

Set-up restated from the script (line numbers of the reference file):
  :77-93    NBSplineKernel order 5, nx1 = 100 nodes on [0, 1], rho = 1, eps = 0, v = -1, nPerh = 1.35, gamma = 5/3
  :98-160   hydroType SPH, cfl 0.5, XSPH False, compatible energy, grad-h correction, corrected velocity gradient, RigorousSumDensity,
            IdealH, hmin 1e-4, hmax 0.1, CheapSynchronousRK2, lastDt 1e-4, dtMin 1e-5, dtMax 0.1, dtGrowth 2
  SPH/SPHHydros.py:84-90   default Q = LimitedMonaghanGingoldViscosity(Cl = 2 (kext/2), Cq = 2 (kext/2)^2), kext = 3 for this kernel
  :600-606  one ReflectingBoundary through x = 0
  SimulationControl/SpheralController.py:231-232, 909-931   iterateIdealH(50 iterations, tolerance 1e-4) at start-up; the derivative
            fields are NOT evaluated at start-up (initializeDerivatives = False), so the first trial advance sees zero derivatives
  :683-685  control.step(5), then control.advance(goalTime = 0.6)
  :786-880  errors against NohAnalyticSolution in the window r in [0.05, 0.35], Pnorm.gridpnorm (SimulationControl/Pnorm.py:58-121)
The stored reference norms (LnormRef["SPH"], :226-240) are in REF below; the script accepts np.allclose(L, ref, rtol = atol = 1e-5).
"""
import numpy as np

import common
from spheral_b200 import nodegen as ng

# Noh-planar-1d.py:226-240 (LnormRef["SPH"]): L1, L2, Linf
REF = {"Mass density": (0.05376370586899846, 0.01472935554844709, 1.6558627223391338),
       "Pressure":     (0.018082144742236444, 0.005431965024057943, 0.6289446614721329),
       "Velocity":     (0.024463871274705278, 0.008419536302504558, 0.8561295316236415),
       "Spec Therm E": (0.010557215425476638, 0.0033659949510588386, 0.355220540682649),
       "h":            (0.00043625606815746957, 0.00012010712699702793, 0.008480811209733824)}
# LnormRef["CRKSPH"] (:241-255), produced with the options of ATS test t200 (:36):
#   --hydroType CRKSPH --cfl 0.25 --KernelConstructor NBSplineKernel --order 7 --nPerh 1.01 --Cl 2.0 --Cq 1.0
REF_CRKSPH = {"Mass density": (0.05064113393844768, 0.015297215762507312, 1.6768360873659973),
              "Pressure":     (0.01687133903296828, 0.006328924998534429, 0.7822574604543725),
              "Velocity":     (0.007746512026971996, 0.0029862099280521903, 0.20321897008372736),
              "Spec Therm E": (0.005051748111924938, 0.0015094950940911932, 0.14418618728403587),
              "h":            (0.00019175169182527455, 6.850786936014129e-05, 0.004376337346566557)}
TOL = 1.0e-5                                    # Noh-planar-1d.py:181


def grid_weights(x, rmin, rmax):
    """Pnorm.computeGridWeighting (SimulationControl/Pnorm.py:58-88)."""
    n = len(x)
    order = np.argsort(x, kind="stable")
    xs = x[order]
    w = np.zeros(n)
    for j in range(n):
        r0 = max(rmin, min(rmax, xs[j] if j == 0 else 0.5*(xs[j - 1] + xs[j])))
        r1 = max(rmin, min(rmax, xs[j] if j == n - 1 else 0.5*(xs[j] + xs[j + 1])))
        w[order[j]] = r1 - r0
    return w


def grid_pnorms(err, x, rmin=0.05, rmax=0.35):
    """Pnorm.gridpnorm for p = 1, 2, "inf" (Pnorm.py:108-121; numpy.linalg.norm of the weighted data)."""
    a = np.abs(err)
    wg = grid_weights(x, rmin, rmax)
    ws = ((x >= rmin) & (x <= rmax)).astype(float)
    return (float(np.sum(wg*a)/max(1e-30, wg.sum())), float(np.sqrt(np.sum((wg*a)**2))/max(1e-30, wg.sum())**0.5),
            float(np.max(ws*a)))


def analytic(t, x, gamma, h0):
    """NohAnalyticSolution.NohSolution(nDim = 1).solution (NohAnalyticSolution.py:27-66), v0 = -1, rho0 = 1."""
    inside = np.abs(x) <= t/3.0
    u = np.where(inside, 0.5, 0.0)
    rho = np.where(inside, 4.0, 1.0)
    return np.where(inside, 0.0, -1.0), u, rho, (gamma - 1.0)*u*rho, np.where(inside, h0/4.0, h0)


def run(orc, first_step_sees_zero_derivatives=True, iterate_initial_H=True, hydro="SPH", volume_policy=True, overrides=None,
        gradhCorrection=True):
    """hydro = "CRKSPH": the set-up of ATS test t200 (NBSpline order 7, nPerh 1.01, cfl 0.25, Cl 2, Cq 1; CRKSPH/CRKSPHHydros.py: LinearOrder
    corrections, LimitedMonaghanGingold Q; controller: RKSumVolume).  volume_policy: CRKSPHBase.cc:155 enrolls the volume with
    ContinuityVolumePolicy (switchable only to show that the golden notices its absence)."""
    crk = hydro == "CRKSPH"
    nx, nPerh, gamma, goal = 100, (1.01 if crk else 1.35), 5.0/3.0, 0.6
    pos, mass, H, d = ng.lattice(1, nx, [0.0], [1.0], 1.0, nPerh)        # distributeNodesInRange1d: x = (i + 1/2) dx, H = 1/(nPerh dx)
    N = nx
    st = dict(position=pos, velocity=-np.ones((N, 1)), H=H, mass=mass, massDensity=np.ones(N), specificThermalEnergy=np.zeros(N),
              pressure=np.zeros(N), soundSpeed=np.zeros(N), omegaGradh=np.ones(N))
    WT = orc.TableKernel(orc.KERNEL_NBSPLINE + (7 if crk else 5), 1, 1000)
    kext = WT.kext
    Cl, Cq = (2.0, 1.0) if crk else (2.0*(kext/2.0), 2.0*(kext/2.0)**2)
    oo = orc.default_options(1, nPerh=nPerh, Qkind=orc.Q_LIMITED_MG, Cl=Cl, Cq=Cq, XSPH=0,
                             compatibleEnergy=1, correctVelocityGradient=1, hmin=1.0e-4, hmax=0.1)
    for k, v in (overrides or {}).items():         # negative controls only
        setattr(oo, k, v)
    so = orc.default_step_options(cfl=0.25 if crk else 0.5)
    rk = common.OracleRK2(orc, oo, so, WT, st, densityUpdate=1, planes=[(np.zeros(1), np.ones(1))], dtMin=1.0e-5, dtMax=0.1,
                          dtGrowth=2.0, crk=crk, volume_policy=volume_policy, gradhCorrection=gradhCorrection)
    rk.s["DvDxQ"] = np.zeros((N, 1))
    if iterate_initial_H:                          # Utilities/iterateIdealH.cc:120-200 for an isotropic ideal H
        done = np.zeros(N, dtype=bool)
        for _ in range(50):
            rk._set_ghosts()
            rk._pairs()
            dd = orc.evaluate_derivatives(oo, WT, rk.s, N, rk.nGhost, rk.pi, rk.pj, rk.cnt)      # the smoothing-scale package alone
            h1 = np.asarray(dd["Hideal"]).reshape(-1)[:N]
            delta = np.abs(h1/rk.s["H"][:N, 0] - 1.0)
            act = ~done
            worst = float(delta[act].max()) if act.any() else 0.0
            done |= act & (delta <= 1.0e-4)
            Hn = rk.s["H"].copy()
            Hn[:N][act, 0] = h1[act]
            rk.s["H"] = Hn
            if worst <= 1.0e-4:
                break
        rk.s = {k: v[:N] for k, v in rk.s.items()}
        rk.nGhost = 0
    rk.lastDt = 1.0e-4                             # integrator.lastDt = dt (:619)
    rk.initializeDerivatives()                     # only to create the derivative fields ...
    if first_step_sees_zero_derivatives:           # ... which the reference leaves at zero until the first evaluation inside step 1
        rk.derivs = {k: np.zeros_like(v) for k, v in rk.derivs.items()}
        rk.pairs_eval = (rk.pi, rk.pj)
    n = 0
    while rk.t < goal:
        rk.step(1.0e100 if n < 5 else goal)        # control.step(5); control.advance(goalTime)
        n += 1
    x = rk.s["pos"][:N, 0]
    rho, eps, v, h = rk.s["rho"][:N], rk.s["eps"][:N], rk.s["vel"][:N, 0], 1.0/rk.s["H"][:N, 0]
    va, ua, rhoa, Pa, ha = analytic(rk.t, x, gamma, 1.0/(1.0/(nPerh*(1.0/nx))))
    out = {}
    for name, data, ans in (("Mass density", rho, rhoa), ("Pressure", (gamma - 1.0)*rho*eps, Pa), ("Velocity", v, va),
                            ("Spec Therm E", eps, ua), ("h", h, ha)):
        out[name] = grid_pnorms(data - ans, x)
    return out, dict(cycles=n, time=rk.t, E=rk.total_energy())
