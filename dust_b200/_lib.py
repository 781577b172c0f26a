"""ctypes binding of libdust_b200.so (the C ABI declared in include/dust_b200.h).

There is NO CPU fallback: if the shared library is missing, or a call is made without a
CUDA device, the import / call raises.  torch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DUST_B200_LIB", os.path.join(_HERE, "libdust_b200.so"))  # env override: A/B builds

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4
MODEL_PENDULUM, MODEL_PARTICLE = 0, 1
PARAMS_BLOCKED, PARAMS_INTERLEAVED = 0, 1
LIK_EXP_UTILITY, LIK_EXPECTED_COST = 0, 1
ROLL_REPEAT, ROLL_MEAN, ROLL_RESAMPLE = 0, 1, 2
SELECT_ARGMAX, SELECT_AVERAGE = 0, 1

_f, _i, _p, _sz = C.c_float, C.c_int32, C.c_void_p, C.c_size_t


class ModelDesc(C.Structure):
    _fields_ = [
        ("kind", _i), ("dt", _f), ("g", _f), ("max_torque", _f), ("max_speed_pend", _f),
        ("w_angle", _f), ("w_speed", _f), ("default_length", _f), ("default_mass", _f),
        ("max_accel", _f), ("max_speed", _f), ("target", _f * 4), ("w_state", _f * 4),
        ("w_term", _f * 4), ("w_ctrl", _f * 2), ("w_obs", _f), ("inv_cell", _f),
        ("c_offset", _f * 2), ("grid_nx", _i), ("grid_ny", _i), ("can_crash", _i),
        ("with_obstacle", _i), ("grid_bits", _p),
    ]


class RolloutArgs(C.Structure):
    _fields_ = [
        ("model", C.POINTER(ModelDesc)), ("B", _i), ("N", _i), ("S", _i), ("P", _i), ("H", _i),
        ("param_tiling", _i), ("likelihood", _i),
        ("state0", _p), ("theta", _p), ("noise", _p), ("sigma", _p), ("params", _p), ("a_seq", _p),
        ("pert", _p), ("alpha", _f), ("temperature", _f),
        ("sigma_weights", _p), ("ctrl_mat", _p), ("ctrl_reg", _f), ("p_begin", _i), ("p_end", _i),
        ("costs", _p), ("log_lik", _p), ("lik_weights", _p), ("grad_lik", _p), ("mppi_weights", _p),
        ("mppi_delta", _p), ("mix", _p), ("states", _p),
        ("workspace", _p), ("workspace_bytes", _sz),
    ]


class SvmpcStepArgs(C.Structure):
    _fields_ = [
        ("rollout", RolloutArgs), ("do_forward", _i), ("roll_strategy", _i), ("weighted_prior", _i), ("prior_aliased", _i),
        ("mu", _p), ("mix", _p), ("inv_var", _p), ("log_norm", _f), ("gamma", _f), ("c1", _f), ("c2", _f), ("lr", _f),
        ("theta_out", _p), ("phi", _p), ("p_weights", _p), ("i_star", _p), ("a_seq", _p), ("theta_next", _p),
        ("mix_next", _p),
    ]


class AdjointArgs(C.Structure):
    _fields_ = [
        ("model", C.POINTER(ModelDesc)), ("B", _i), ("N", _i), ("S", _i), ("P", _i), ("H", _i),
        ("param_tiling", _i), ("likelihood", _i),
        ("state0", _p), ("theta", _p), ("noise", _p), ("sigma", _p), ("params", _p), ("lik_weights", _p),
        ("alpha", _f), ("grad_theta", _p), ("grad_params", _p),
        ("workspace", _p), ("workspace_bytes", _sz), ("p_begin", _i), ("p_end", _i),
    ]


class GmmArgs(C.Structure):
    _fields_ = [
        ("B", _i), ("M", _i), ("K", _i), ("D", _i), ("x", _p), ("mu", _p), ("mix", _p),
        ("inv_var", _p), ("log_norm", _f), ("log_prob", _p), ("score", _p),
    ]


class MedianArgs(C.Structure):
    _fields_ = [
        ("N", _i), ("D", _i), ("row_begin", _i), ("row_end", _i), ("x", _p), ("hist", _p),
        ("selected", _p), ("row_norms", _p), ("sample_begin", _i), ("sample_end", _i), ("ld", _i),
    ]


class PeerArgs(C.Structure):
    _fields_ = [
        ("world", _i), ("rank", _i), ("epoch", _i), ("rows_per_rank", _i), ("row_floats", _i),
        ("slabs", C.POINTER(_p)), ("flags", C.POINTER(_p)), ("gathered", _p),
        ("gathered_peers", C.POINTER(_p)), ("counters", _p), ("n_parts", _i), ("part_floats", _i * 4), ("parts", _p * 4),
    ]


class PhiArgs(C.Structure):
    _fields_ = [
        ("B", _i), ("N", _i), ("D", _i), ("row_begin", _i), ("row_end", _i), ("per_dim", _i),
        ("x", _p), ("score", _p), ("gamma", _f), ("c1", _f), ("c2", _f), ("gamma_dev", _p),
        ("bw_scale", _f), ("lr", _f), ("phi", _p), ("x_out", _p), ("bandwidths", _p),
        ("workspace", _p), ("workspace_bytes", _sz), ("x_prepared", _i), ("ld", _i),
    ]


class SvmpcForwardArgs(C.Structure):
    _fields_ = [
        ("B", _i), ("N", _i), ("H", _i), ("A", _i), ("roll_strategy", _i), ("weighted_prior", _i),
        ("log_lik", _p), ("theta", _p), ("mu", _p), ("mix", _p), ("inv_var", _p), ("log_norm", _f),
        ("p_weights", _p), ("i_star", _p), ("a_seq", _p), ("theta_next", _p), ("mix_next", _p),
        ("resample_noise", _p),
    ]


class DiscoStepArgs(C.Structure):
    _fields_ = [
        ("B", _i), ("N", _i), ("H", _i), ("A", _i), ("strategy", _i), ("steps", _i),
        ("a_low", _p), ("a_high", _p), ("a_mat", _p), ("a_mix", _p), ("a_seq", _p),
        ("next_actions", _p),
    ]


class MpfArgs(C.Structure):
    _fields_ = [
        ("model", C.POINTER(ModelDesc)), ("B", _i), ("Np", _i), ("n_steps", _i), ("log_space", _i),
        ("x", _p), ("obs0", _p), ("action", _p), ("obs1", _p), ("prior_inv_var", _p),
        ("obs_std", _f), ("bw", _f), ("lr", _f), ("grad_norms", _p),
        ("workspace", _p), ("workspace_bytes", C.c_size_t), ("bw_dev", _p), ("phi_out", _p),
    ]


# every symbol include/dust_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "dust_rollout_workspace_bytes": (_sz, [C.POINTER(RolloutArgs)]),
    "dust_rollout_cost": (C.c_int, [C.POINTER(RolloutArgs), _p]),
    "dust_rollout_plan": (C.c_int, [C.POINTER(RolloutArgs), C.POINTER(_i * 5)]),
    "dust_cost_reduce": (C.c_int, [C.POINTER(RolloutArgs), _p]),
    "dust_phi_tc_plan": (C.c_int, [_i, _i, C.POINTER(_i * 22)]),
    "dust_phi_tc_mode": (C.c_int, [_i]),
    "dust_peer_alloc": (C.c_int, [_sz, C.POINTER(_p)]),
    "dust_peer_free": (C.c_int, [_p]),
    "dust_peer_export": (C.c_int, [_p, C.POINTER(C.c_ubyte * 64)]),
    "dust_peer_open": (C.c_int, [C.POINTER(C.c_ubyte * 64), C.POINTER(_p)]),
    "dust_peer_close": (C.c_int, [_p]),
    "dust_peer_signal": (C.c_int, [C.POINTER(PeerArgs), _p]),
    "dust_peer_gather": (C.c_int, [C.POINTER(PeerArgs), _p]),
    "dust_peer_push": (C.c_int, [C.POINTER(PeerArgs), _p]),
    "dust_peer_wait": (C.c_int, [C.POINTER(PeerArgs), _p]),
    "dust_svmpc_step": (C.c_int, [C.POINTER(SvmpcStepArgs), _p]),
    "dust_adjoint_workspace_bytes": (_sz, [C.POINTER(AdjointArgs)]),
    "dust_rollout_adjoint": (C.c_int, [C.POINTER(AdjointArgs), _p]),
    "dust_adjoint_plan": (C.c_int, [C.POINTER(AdjointArgs), C.POINTER(_i * 5)]),
    "dust_gmm_score": (C.c_int, [C.POINTER(GmmArgs), _p]),
    "dust_median_hist_pass": (C.c_int, [C.POINTER(MedianArgs), _i, _p]),
    "dust_median_select": (C.c_int, [C.POINTER(MedianArgs), _i, _p, _p]),
    "dust_median_fast_supported": (C.c_int, [_i, _i]),
    "dust_median_fast_workspace_bytes": (_sz, [_i, _i]),
    "dust_median_fast_prepare": (C.c_int, [C.POINTER(MedianArgs), _p, _sz, _p]),
    "dust_median_fast_sample_hist_offset": (_sz, [_i, _i]),
    "dust_median_fast_window": (C.c_int, [C.POINTER(MedianArgs), _p, _sz, _p]),
    "dust_median_fast_count": (C.c_int, [C.POINTER(MedianArgs), _p, _sz, _p]),
    "dust_median_fast_select": (C.c_int, [C.POINTER(MedianArgs), _p, _p]),
    "dust_phi_workspace_bytes": (_sz, [C.POINTER(PhiArgs)]),
    "dust_svgd_phi": (C.c_int, [C.POINTER(PhiArgs), _p]),
    "dust_bandwidth_from_median": (C.c_int, [_p, _i, _f, _i, _p, _p]),
    "dust_svmpc_forward": (C.c_int, [C.POINTER(SvmpcForwardArgs), _p]),
    "dust_disco_step": (C.c_int, [C.POINTER(DiscoStepArgs), _p]),
    "dust_mpf_workspace_bytes": (C.c_size_t, [C.POINTER(MpfArgs)]),
    "dust_mpf_optimize": (C.c_int, [C.POINTER(MpfArgs), _p]),
    "dust_silverman_bandwidth": (C.c_int, [_p, _i, _f, _p, _p, _i, _p]),
    "dust_model_step": (C.c_int, [C.POINTER(ModelDesc), _i, _p, _p, _p, _p, _p]),
    "dust_model_cost": (C.c_int, [C.POINTER(ModelDesc), _i, _i, _p, _p, _p, _p]),
    "dust_aux_model_step": (C.c_int, [_i, _f, C.POINTER(_f * 8), _i, _p, _p, _p, _p, _p]),
    "dust_noise_normal": (C.c_int, [_p, C.c_int64, C.c_uint64, C.c_uint64, _p]),
    "dust_profiler_enable": (None, [C.c_int]),
    "dust_profiler_reset": (None, []),
    "dust_profiler_report": (C.c_int, [C.c_char_p, _sz]),
    "dust_launch_count": (C.c_ulonglong, []),
    "dust_abi_version": (C.c_int, []),
    "dust_last_error": (C.c_char_p, []),
    "dust_build_info": (C.c_char_p, []),
}

_lib = None
launch_count = 0  # kernels-launching ABI calls made through this module (bench.py reads it)


def load():
    """dlopen libdust_b200.so (once) and declare the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C dust_b200/csrc`).  dust_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.dust_abi_version() != 3:
        raise ImportError("libdust_b200.so ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


_EXC = {ERR_INVALID_ARG: ValueError, ERR_UNSUPPORTED: NotImplementedError,
        ERR_WORKSPACE: RuntimeError, ERR_CUDA: RuntimeError}


def check(rc):
    if rc != OK:
        msg = load().dust_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, RuntimeError)(msg)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dust_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t):
    """device pointer of a contiguous float32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("dust_b200: tensor must live on a CUDA device")
    if not t.is_contiguous():
        raise ValueError("dust_b200: tensor must be contiguous")
    if t.dtype not in (torch.float32, torch.int32):
        raise ValueError(f"dust_b200: tensor must be float32 (int32 for indices), got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"dust_b200: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                         "(kernels are enqueued on the current device's stream: wrap the call in torch.cuda.device(...))")
    return t.data_ptr()


def ptr_rows(t):
    """(device pointer, row stride in floats) of a 2-D float32 CUDA tensor whose rows are contiguous but may be spaced
    (a column slice of a wider buffer, e.g. X = packed[:, :D])."""
    if t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda or t.dtype != torch.float32:
        raise ValueError("dust_b200: expected a 2-D float32 CUDA tensor with unit column stride")
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"dust_b200: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}")
    return t.data_ptr(), int(t.stride(0))


def profiler_report():
    """{kernel name: (launches, total_ms)} accumulated since dust_profiler_reset (synchronises)."""
    buf = C.create_string_buffer(1 << 16)
    check(load().dust_profiler_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(" ", 2)
        out[name] = (int(n), float(ms))
    return out


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args, launches=1):
    global launch_count
    launch_count += launches
    check(getattr(load(), name)(*args))
