"""ctypes binding of libsphb200.so (the C ABI in include/sphb200.h).

There is no CPU fallback: if the shared library is missing, importing this module's `lib()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SPHB200_LIB: A/B measurements of tuning variants of the same library (scripts/build_variants.py); never a fallback
LIB_PATH = os.environ.get("SPHB200_LIB") or os.path.join(HERE, "libsphb200.so")

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)

Q_MG, Q_LIMITED_MG = 0, 1
H_SPH, H_ASPH, H_NONE, H_ASPH_CLASSIC = 0, 1, 2, 3
KERNEL_BSPLINE, KERNEL_WENDLANDC4, KERNEL_WENDLANDC2 = 0, 1, 2
KERNEL_NBSPLINE = 100        # + order: NBSplineKernel(order)
TABLE_W, TABLE_WPI = 0, 1
HYDRO_SPH, HYDRO_CRKSPH = 0, 1

STATE_FIELDS = ("position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy",
                "pressure", "soundSpeed", "omegaGradh", "DvDxQ", "fCl", "fCq", "volume", "rkCorrections")
STATE_BITS = {k: 1 << i for i, k in enumerate(STATE_FIELDS)}
DERIV_FIELDS = ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "localDvDx", "gradRho", "M", "localM",
                "rhoSum", "normalization", "maxViscousPressure", "effViscousPressure", "XSPHWeightSum",
                "XSPHDeltaV", "DHDt", "Hideal", "massZerothMoment", "massFirstMoment")
DERIV_BITS = {k: 1 << i for i, k in enumerate(DERIV_FIELDS)}


def state_width(ndim, name):
    if name in ("position", "velocity"):
        return ndim
    if name == "H":
        return 6 if ndim == 3 else 3
    if name == "DvDxQ":
        return ndim*ndim
    if name == "rkCorrections":
        return (ndim + 1)*(ndim + 1)
    return 1


def deriv_width(ndim, name):
    if name in ("DxDt", "DvDt", "gradRho", "XSPHDeltaV", "massFirstMoment"):
        return ndim
    if name in ("DvDx", "localDvDx", "M", "localM"):
        return ndim*ndim
    if name in ("DHDt", "Hideal"):
        return 6 if ndim == 3 else 3
    return 1


class Options(C.Structure):
    _fields_ = [("ndim", C.c_int), ("compatibleEnergy", C.c_int), ("evolveTotalEnergy", C.c_int),
                ("XSPH", C.c_int), ("correctVelocityGradient", C.c_int),
                ("epsTensile", C.c_double), ("nTensile", C.c_double), ("nPerh", C.c_double),
                ("Qkind", C.c_int), ("Cl", C.c_double), ("Cq", C.c_double), ("eps2", C.c_double),
                ("negligibleSoundSpeed", C.c_double),
                ("balsara", C.c_int), ("linearInExpansion", C.c_int), ("quadraticInExpansion", C.c_int),
                ("etaCritFrac", C.c_double), ("etaFoldFrac", C.c_double),
                ("hEvolution", C.c_int), ("hmin", C.c_double), ("hmax", C.c_double), ("hydro", C.c_int), ("hminratio", C.c_double)]


class HostState(C.Structure):
    _fields_ = [(k, _dp) for k in STATE_FIELDS]


class HostDerivs(C.Structure):
    _fields_ = [(k, _dp) for k in DERIV_FIELDS]


class GammaLaw(C.Structure):
    _fields_ = [("gamma", C.c_double), ("minimumPressure", C.c_double), ("maximumPressure", C.c_double),
                ("externalPressure", C.c_double), ("minPressureType", C.c_int)]


HEVOLUTION_IDEALH, HEVOLUTION_INTEGRATEH, HEVOLUTION_FIXEDH = 0, 1, 2
DT_REASONS = ("sound speed", "artificial viscosity", "velocity divergence", "acceleration", "velocity magnitude",
              "pairwise velocity difference")


class StepOptions(C.Structure):
    _fields_ = [("eos", GammaLaw), ("rhoMin", C.c_double), ("rhoMax", C.c_double), ("hminratio", C.c_double),
                ("HEvolution", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("ms_build_pairs", C.c_float), ("ms_evaluate", C.c_float),
                ("ms_energy", C.c_float), ("ms_pair_kernel", C.c_float), ("ms_neighbor_kernels", C.c_float),
                ("directed_edges", C.c_uint64), ("stencil_radius", C.c_uint32), ("fine_walk", C.c_uint32)]


EXPORTS = ("sphb200_abi_version", "sphb200_create", "sphb200_destroy", "sphb200_set_options", "sphb200_last_error",
           "sphb200_sync", "sphb200_set_kernel_table", "sphb200_table_ncoef", "sphb200_table_kernel_build",
           "sphb200_set_nodes", "sphb200_upload_state", "sphb200_upload_state_values", "sphb200_download_state", "sphb200_build_pairs",
           "sphb200_download_pairs", "sphb200_download_neighbor_counts", "sphb200_evaluate_derivatives",
           "sphb200_download_derivs", "sphb200_download_pair_accelerations", "sphb200_copy_DvDx_to_Q",
           "sphb200_update_energy_compatible", "sphb200_halo_bytes_per_node", "sphb200_halo_pack",
           "sphb200_halo_unpack", "sphb200_node_bounds", "sphb200_halo_select", "sphb200_stream", "sphb200_get_stats", "sphb200_measure_fp64_peak",
           "sphb200_node_bounds_device", "sphb200_halo_select_device",
           "sphb200_crk_compute_volume", "sphb200_crk_compute_corrections", "sphb200_crk_sum_mass_density",
           "sphb200_sum_mass_density", "sphb200_compute_omega_gradh", "sphb200_update_eos_gamma_law", "sphb200_state_copy",
           "sphb200_state_assign", "sphb200_state_update", "sphb200_compute_dt",
           "sphb200_reflect_configure", "sphb200_reflect_set_ghost_nodes", "sphb200_reflect_apply_ghosts", "sphb200_reflect_enforce",
           "sphb200_reflect_finalize_derivatives", "sphb200_halo_unpack_values", "sphb200_halo_pack_derivs",
           "sphb200_halo_unpack_derivs", "sphb200_upload_derivs", "sphb200_boundary_configure", "sphb200_iterate_ideal_h", "sphb200_connectivity_valid", "sphb200_state_fields_present",
           "sphb200_evaluate_derivatives_to_host")

_lib = None


def lib():
    """Load libsphb200.so; fail loudly if it has not been built (python -m spheral_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsphb200.so is missing at %s -- build it with `python -m spheral_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.sphb200_abi_version.restype = C.c_int
    L.sphb200_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Options)]
    L.sphb200_destroy.argtypes = [vp]
    L.sphb200_destroy.restype = None
    L.sphb200_set_options.argtypes = [vp, C.POINTER(Options)]
    L.sphb200_last_error.argtypes = [vp]
    L.sphb200_last_error.restype = C.c_char_p
    L.sphb200_sync.argtypes = [vp]
    L.sphb200_set_kernel_table.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_size_t, _dp, _dp, _dp,
                                           C.c_size_t, C.c_double, C.c_double, _dp,
                                           C.c_size_t, C.c_double, C.c_double, _dp]
    L.sphb200_table_ncoef.argtypes = [C.c_size_t]
    L.sphb200_table_ncoef.restype = C.c_size_t
    L.sphb200_table_kernel_build.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_double,
                                             _dp, _dp, C.POINTER(C.c_size_t), _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    L.sphb200_set_nodes.argtypes = [vp, C.c_size_t, C.c_size_t]
    L.sphb200_upload_state.argtypes = [vp, C.c_uint, C.POINTER(HostState)]
    L.sphb200_upload_state_values.argtypes = [vp, C.c_uint, C.POINTER(HostState)]
    L.sphb200_download_state.argtypes = [vp, C.c_uint, C.POINTER(_dp)]
    L.sphb200_build_pairs.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sphb200_download_pairs.argtypes = [vp, _u32p, _u32p, C.c_size_t]
    L.sphb200_download_neighbor_counts.argtypes = [vp, _u32p]
    L.sphb200_evaluate_derivatives.argtypes = [vp, C.c_double, C.c_double]
    L.sphb200_download_derivs.argtypes = [vp, C.c_uint, C.POINTER(HostDerivs)]
    L.sphb200_evaluate_derivatives_to_host.argtypes = [vp, C.c_double, C.c_double, C.c_uint, C.POINTER(HostDerivs)]
    L.sphb200_download_pair_accelerations.argtypes = [vp, _dp, C.c_size_t]
    L.sphb200_copy_DvDx_to_Q.argtypes = [vp]
    L.sphb200_update_energy_compatible.argtypes = [vp, C.c_double]
    L.sphb200_halo_bytes_per_node.argtypes = [vp, C.c_uint]
    L.sphb200_halo_bytes_per_node.restype = C.c_size_t
    L.sphb200_halo_pack.argtypes = [vp, C.c_uint, vp, C.c_size_t, vp]
    L.sphb200_halo_unpack.argtypes = [vp, C.c_uint, C.c_size_t, C.c_size_t, vp]
    L.sphb200_node_bounds.argtypes = [vp, C.c_size_t, _dp, _dp, _dp]
    L.sphb200_halo_select.argtypes = [vp, C.c_int, C.c_size_t, C.c_double, C.c_double, C.c_double,
                                      vp, C.POINTER(C.c_size_t), vp, C.POINTER(C.c_size_t), C.c_size_t]
    L.sphb200_node_bounds_device.argtypes = [vp, C.c_size_t, vp]
    L.sphb200_halo_select_device.argtypes = [vp, C.c_int, C.c_size_t, C.c_double, C.c_double, vp, vp, vp, vp, C.c_size_t]
    L.sphb200_stream.argtypes = [vp]
    L.sphb200_stream.restype = vp
    L.sphb200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.sphb200_measure_fp64_peak.argtypes = [vp, _dp]
    L.sphb200_crk_compute_volume.argtypes = [vp]
    L.sphb200_crk_compute_corrections.argtypes = [vp]
    L.sphb200_crk_sum_mass_density.argtypes = [vp, C.c_double, C.c_double]
    L.sphb200_sum_mass_density.argtypes = [vp]
    L.sphb200_compute_omega_gradh.argtypes = [vp]
    L.sphb200_update_eos_gamma_law.argtypes = [vp, C.POINTER(GammaLaw)]
    L.sphb200_state_copy.argtypes = [vp]
    L.sphb200_state_assign.argtypes = [vp]
    L.sphb200_state_update.argtypes = [vp, C.POINTER(StepOptions), C.c_double, C.c_int]
    L.sphb200_compute_dt.argtypes = [vp, C.c_double, C.c_int, _dp, C.POINTER(C.c_int), _u32p]
    L.sphb200_reflect_configure.argtypes = [vp, C.c_int, _dp, _dp]
    L.sphb200_reflect_set_ghost_nodes.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sphb200_reflect_apply_ghosts.argtypes = [vp, C.c_uint]
    L.sphb200_reflect_enforce.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sphb200_reflect_finalize_derivatives.argtypes = [vp]
    L.sphb200_halo_unpack_values.argtypes = [vp, C.c_uint, C.c_size_t, C.c_size_t, vp]
    L.sphb200_halo_pack_derivs.argtypes = [vp, vp, C.c_size_t, vp]
    L.sphb200_halo_unpack_derivs.argtypes = [vp, C.c_size_t, C.c_size_t, vp]
    L.sphb200_upload_derivs.argtypes = [vp, C.c_uint, C.POINTER(HostDerivs)]
    L.sphb200_boundary_configure.argtypes = [vp, C.c_int, C.POINTER(C.c_int), _dp, _dp, _dp, _dp]
    L.sphb200_iterate_ideal_h.argtypes = [vp, C.c_int, C.c_double, _dp]
    L.sphb200_connectivity_valid.argtypes = [vp]
    L.sphb200_state_fields_present.argtypes = [vp]
    L.sphb200_state_fields_present.restype = C.c_uint
    _lib = L
    return L
