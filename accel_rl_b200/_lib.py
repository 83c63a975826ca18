"""ctypes binding of libaccelrl_b200.so (include/accelrl_b200.h).

The product path has NO CPU fallback: importing this module never touches the oracle, and
every compute entry point raises if the shared library is missing or no sm_100a device exists.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("ARL_LIB_PATH", os.path.join(CSRC, "libaccelrl_b200.so"))   # override: debug builds only

ARL_MAX_CONV = 4
IPC_HANDLE_BYTES = 64


class NetCfg(C.Structure):
    _fields_ = [
        ("n_conv", C.c_int),
        ("conv_filters", C.c_int * ARL_MAX_CONV),
        ("conv_sizes", C.c_int * ARL_MAX_CONV),
        ("conv_strides", C.c_int * ARL_MAX_CONV),
        ("conv_pads", C.c_int * ARL_MAX_CONV),
        ("hidden", C.c_int),
        ("n_actions", C.c_int),
        ("in_c", C.c_int),
        ("in_h", C.c_int),
        ("in_w", C.c_int),
        ("pixel_scale", C.c_float),
        ("max_rows", C.c_int),
    ]


class OptCfg(C.Structure):
    _fields_ = [
        ("algo", C.c_int),
        ("clip_param", C.c_float),
        ("v_loss_coeff", C.c_float),
        ("ent_loss_coeff", C.c_float),
        ("update", C.c_int),
        ("learning_rate", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("epsilon", C.c_float),
        ("rho", C.c_float),
        ("grad_norm_clip", C.c_float),
        ("ppo_tie_grad", C.c_int),
    ]


class SamplerCfg(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int),
        ("horizon", C.c_int),
        ("planes", C.c_int),
        ("observations", C.c_void_p),
        ("rewards", C.c_void_p),
        ("dones", C.c_void_p),
        ("raw_reward", C.c_void_p),
        ("need_reset", C.c_void_p),
        ("actions", C.c_void_p),
        ("prob", C.c_void_p),
        ("value", C.c_void_p),
        ("extra_observations", C.c_void_p),
        ("step_obs", C.c_void_p),
        ("uniforms", C.c_void_p),
        ("frame_pool", C.c_void_p),
        ("pool_frames", C.c_int),
        ("max_path_length", C.c_int),
        ("discount", C.c_float),
        ("mid_batch_reset", C.c_int),
        ("clip_reward", C.c_int),
        ("episodic_lives", C.c_int),
        ("lives0", C.c_int),
        ("life_base", C.c_int),
        ("life_mul", C.c_int),
        ("life_mod", C.c_int),
        ("reward_mod", C.c_int),
        ("frame_stride", C.c_int),
        ("traj_cap", C.c_int),
        ("ext_emulator", C.c_int),
        ("n_games", C.c_int),
        ("frame_mode", C.c_int),
    ]


class ExtStep(C.Structure):
    """arl_ext_step"""
    _fields_ = [("reward", C.c_float), ("raw_reward", C.c_float), ("done", C.c_uint8), ("need_reset", C.c_uint8),
                ("flags", C.c_uint8), ("pad", C.c_uint8)]


# name -> (restype, argtypes); must list every symbol include/accelrl_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "arl_create": (C.c_int, [C.POINTER(NetCfg), C.POINTER(_P)]),
    "arl_destroy": (None, [_P]),
    "arl_last_error": (C.c_char_p, [_P]),
    "arl_device_error": (C.c_int, [_P]),
    "arl_param_count": (C.c_long, [_P]),
    "arl_param_layout": (C.c_int, [_P, C.POINTER(C.c_long), C.POINTER(C.c_long), C.c_int]),
    "arl_bind_params": (C.c_int, [_P, _P, _P, _P, _P]),
    "arl_pack_weights": (C.c_int, [_P, _P]),
    "arl_policy_forward": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "arl_sample_actions": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "arl_frame_update": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "arl_frame_update_rgb": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "arl_sampler_configure": (C.c_int, [_P, C.POINTER(SamplerCfg)]),
    "arl_sampler_select": (C.c_int, [_P, C.c_int]),
    "arl_sampler_reset": (C.c_int, [_P, _P]),
    "arl_sampler_warmup": (C.c_int, [_P, _P, C.c_int, _P]),
    "arl_rollout_begin": (C.c_int, [_P, _P]),
    "arl_rollout_step": (C.c_int, [_P, C.c_int, _P, _P]),
    "arl_rollout_end": (C.c_int, [_P, _P]),
    "arl_rollout_run": (C.c_int, [_P, _P]),
    "arl_rollout_serve": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "arl_rollout_ingest": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "arl_host_register": (C.c_int, [_P, C.c_size_t]),
    "arl_host_unregister": (C.c_int, [_P]),
    "arl_copy_async": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, _P]),
    "arl_traj_read": (C.c_int, [_P, C.POINTER(C.c_int), _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "arl_peek_frame_cmds": (C.c_int, [_P, _P, C.c_int, _P]),
    "arl_gae": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "arl_opt_configure": (C.c_int, [_P, C.POINTER(OptCfg)]),
    "arl_bind_train_inputs": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_long]),
    "arl_set_lr_mult": (C.c_int, [_P, C.c_float, _P]),
    "arl_grad_minibatch": (C.c_int, [_P, _P, C.c_int, _P]),
    "arl_clip_update": (C.c_int, [_P, C.c_float, _P]),
    "arl_train_minibatches": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "arl_train_minibatches_sync": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "arl_train_minibatches_async": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "arl_read_logs": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_int), _P]),
    "arl_reset_opt_state": (C.c_int, [_P, _P]),
    "arl_opt_step_get": (C.c_int, [_P, C.POINTER(C.c_int), _P]),
    "arl_opt_step_set": (C.c_int, [_P, C.c_int, _P]),
    "arl_comm_local_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "arl_comm_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "arl_comm_connect": (C.c_int, [_P, _P]),
    "arl_sync_allreduce_update": (C.c_int, [_P, _P]),
    "arl_comm_trace": (C.c_int, [_P, _P, C.c_int, _P]),
    "arl_comm_barrier": (C.c_int, [_P, _P]),
    "arl_async_local_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "arl_async_connect": (C.c_int, [_P, _P]),
    "arl_async_regions": (C.c_int, [_P]),
    "arl_async_push_pull": (C.c_int, [_P, _P]),
    "arl_async_pull": (C.c_int, [_P, _P]),
    "arl_async_read_central": (C.c_int, [_P, C.c_int, _P, C.c_long, _P]),
    "arl_debug_activation": (C.c_int, [_P, C.c_int, _P, C.c_long, C.POINTER(C.c_long), _P]),
    "arl_kernel_launches": (C.c_long, [_P]),
    "arl_source_hash": (C.c_char_p, []),
    "arl_profile_begin": (C.c_int, [_P, _P]),
    "arl_profile_timeline": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_char_p, C.c_int, _P, C.c_int, _P, _P]),
    "arl_profile_graph": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_char_p, C.c_int, _P, C.c_int,
                                    C.POINTER(C.c_int), _P]),
    "arl_profile_end": (C.c_int, [_P, C.c_char_p, C.c_int, _P, C.c_int, C.POINTER(C.c_int), _P]),
    "arl_test_gemm": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "arl_test_wgrad": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
}

_lib = None


def build(verbose=False):
    """Compile libaccelrl_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libaccelrl_b200.so failed")
    return LIB_PATH


_HASH_FILES = ["api.cu", "common.cuh", "gemm_tc.cuh", "pconv.cuh", "fcgemm.cuh", "kernels.cuh", "comm.cuh",
               os.path.join("..", "..", "include", "accelrl_b200.h")]          # = $(SRC) $(HDR) of csrc/Makefile


def source_hash():
    """sha256 of the CUDA sources in the tree, computed the way csrc/Makefile does"""
    import hashlib
    h = hashlib.sha256()
    for f in _HASH_FILES:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def check_source_hash():
    """raise if the loaded library was not built from the sources next to it"""
    lib = load()
    built = lib.arl_source_hash().decode()
    here = source_hash()
    if built != here:
        raise RuntimeError("libaccelrl_b200.so was built from other sources (library %s..., tree %s...): rebuild with "
                           "`python -c 'import __graft_entry__ as g; g.build()'`" % (built[:12], here[:12]))
    return built


def load():
    """Load the shared library and type every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libaccelrl_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ArlError(RuntimeError):
    pass


def check(ctx, rc):
    if rc != 0:
        msg = load().arl_last_error(ctx)
        raise ArlError("libaccelrl_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
