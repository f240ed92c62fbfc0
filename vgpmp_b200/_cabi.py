"""ctypes binding of include/vgpmp_b200.h.  There is no CPU fallback: a missing library is a hard error."""
import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libvgpmp_b200.so"

MAX_DOF, MAX_SPHERES, MAX_MP = 8, 64, 32
c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class RobotDesc(C.Structure):
    _fields_ = [("dof", C.c_int32), ("craig", C.c_int32), ("num_spheres", C.c_int32),
                ("dh", c_double_p), ("twist", c_double_p), ("base_pose", c_double_p),
                ("sphere_frame", c_int32_p), ("sphere_offsets", c_double_p), ("sphere_radii", c_double_p),
                ("limits_lo", c_double_p), ("limits_hi", c_double_p)]


class SdfDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("data", c_double_p),
                ("origin", C.c_double * 3), ("delta", C.c_double)]


class LikDesc(C.Structure):
    _fields_ = [("sigma_obs", C.c_double), ("epsilon", C.c_double), ("alpha", C.c_double),
                ("scene_offset", C.c_double * 3), ("jitter", C.c_double)]


class Dims(C.Structure):
    _fields_ = [("num_problems", C.c_int32), ("num_inducing", C.c_int32), ("num_timesteps", C.c_int32),
                ("num_samples", C.c_int32), ("num_bases", C.c_int32), ("total_samples", C.c_int32),
                ("kl_shards", C.c_int32)]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("q_mu", "q_sqrt", "lengthscales", "variances", "query_latent", "Z", "X")]


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_q_mu", "d_q_sqrt", "d_lengthscales", "d_variances")]


class Draws(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("omega", "tau", "w", "eps_u", "eps_j")]


class Aux(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("f", "logp", "kl")]


class Adam(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("q_mu", "q_sqrt", "raw_lengthscales", "raw_variances", "lengthscales",
                                          "variances", "m", "v")] + \
               [("variance_lower", C.c_double), ("learning_rate", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("step", C.c_int32), ("train_q_mu", C.c_int32),
                ("train_q_sqrt", C.c_int32), ("train_lengthscales", C.c_int32), ("train_variances", C.c_int32)]


# every symbol include/vgpmp_b200.h declares: name -> (restype, argtypes)
_P, _I, _I64, _U64, _D, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_size_t
SYMBOLS = {
    "vgpmp_create": (_I, [C.POINTER(_P), _I, C.POINTER(RobotDesc), C.POINTER(SdfDesc), C.POINTER(LikDesc)]),
    "vgpmp_create_shared": (_I, [C.POINTER(_P), _P, C.POINTER(RobotDesc), C.POINTER(LikDesc)]),
    "vgpmp_sdf_records_id": (_U64, [_P]),
    "vgpmp_destroy": (_I, [_P]),
    "vgpmp_last_error": (C.c_char_p, [_P]),
    "vgpmp_version": (C.c_char_p, []),
    "vgpmp_workspace_bytes": (_SZ, [_P, C.POINTER(Dims)]),
    "vgpmp_launch_count": (_U64, [_P]),
    "vgpmp_fk_frames": (_I, [_P, _P, _P, _I64, _P]),
    "vgpmp_fk_spheres": (_I, [_P, _P, _P, _I64, _P]),
    "vgpmp_sdf_lookup": (_I, [_P, _P, _P, _P, _I64, _P]),
    "vgpmp_loglik_fwd_bwd": (_I, [_P, _P, _I, _D, _P, _P, _I64, _P]),
    "vgpmp_clearance": (_I, [_P, _P, _P, _I64, _P]),
    "vgpmp_predict_f_mean": (_I, [_P, C.POINTER(Dims), C.POINTER(Params), _P, _I, _P, _P, _SZ, _P]),
    "vgpmp_kuu": (_I, [_P, _P, _P, _P, _D, _P, _I, _I, _P]),
    "vgpmp_kuf": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "vgpmp_gp_prepare": (_I, [_P, C.POINTER(Dims), C.POINTER(Params), _P, _P, _P, _P, _SZ, _P]),
    "vgpmp_pathwise_sample": (_I, [_P, C.POINTER(Dims), C.POINTER(Params), C.POINTER(Draws), _P, _I, _P, _P, _SZ, _P]),
    "vgpmp_elbo_fwd_bwd": (_I, [_P, C.POINTER(Dims), C.POINTER(Params), C.POINTER(Draws), _P, C.POINTER(Grads),
                                C.POINTER(Aux), _P, _SZ, _P]),
    "vgpmp_adam_step": (_I, [_P, C.POINTER(Dims), C.POINTER(Adam), C.POINTER(Grads), _P]),
    "vgpmp_rng_fill": (_I, [_P, C.POINTER(Dims), _U64, _U64, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "vgpmp_rng_fill_lazy": (_I, [_P, C.POINTER(Dims), _U64, _U64, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "vgpmp_rng_fill_async": (_I, [_P, C.POINTER(Dims), _U64, _U64, _I64, _I64, _P, _P, _P, _P, _P, _I]),
    "vgpmp_rng_join": (_I, [_P, _I, _P]),
    "vgpmp_rng_release": (_I, [_P, _I, _P]),
    "vgpmp_train_step_host": (_I, [_P, C.POINTER(Dims), C.POINTER(Adam), _P, _P, _P, _P, _U64, _P, _SZ,
                                   C.POINTER(Grads), _P, _P, _P, _SZ, _P]),
    "vgpmp_train_step_host_begin": (_I, [_P, C.POINTER(Dims), C.POINTER(Adam), _P, _P, _P, _P, _U64, _I64, _P, _SZ,
                                         C.POINTER(Grads), _P, _P, _P, _SZ, _P]),
    "vgpmp_train_step_host_end": (_I, [_P, C.POINTER(Dims), _P, _P]),
    "vgpmp_draws_bytes": (_SZ, [C.POINTER(Dims), _I]),
    "vgpmp_draws_bytes_lazy": (_SZ, [C.POINTER(Dims), _I]),
    "vgpmp_sampler_generates_draws": (_I, [_P, C.POINTER(Dims)]),
    "vgpmp_mesh_to_sdf": (_I, [_I, c_double_p, c_double_p, c_int32_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, c_double_p, _D, c_double_p]),
    "vgpmp_probe_fp64_tflops": (_D, [_I]),
    "vgpmp_set_option": (_I, [_P, C.c_char_p, _I]),
    "vgpmp_profile_enable": (_I, [_P, _I]),
    "vgpmp_profile_collect": (_I, [_P, c_double_p, C.POINTER(C.c_int64)]),
    "vgpmp_stage_name": (C.c_char_p, [_I]),
}
NUM_STAGES = 7

_lib = None


class VgpmpError(RuntimeError):
    pass


def load():
    """dlopen libvgpmp_b200.so and type every entry point.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("VGPMP_B200_LIB", LIB_PATH))
    if not path.exists():
        raise VgpmpError(f"{path} is missing: build it with `python -m vgpmp_b200.build` "
                         "(vgpmp_b200 has no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(handle, rc, what=""):
    if rc != 0:
        msg = load().vgpmp_last_error(handle)
        raise VgpmpError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
