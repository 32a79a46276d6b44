"""ctypes binding of libezb200.so (include/ezb200.h).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libezb200.so")


class EzbError(RuntimeError):
    pass


class DitDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("embed_dim", "num_heads", "depth", "context_dim", "inner_dim", "ada_rank")] + \
               [("ada_scaling", C.c_float)] + \
               [(n, C.c_int32) for n in ("latent_chans", "is_controlnet", "cond_c0", "cond_c1", "max_batch", "max_len",
                                         "max_ctx_len", "max_timesteps", "precision")]


class VaeDesc(C.Structure):
    _fields_ = [("latent_dim", C.c_int32), ("channels", C.c_int32), ("out_channels", C.c_int32), ("n_stages", C.c_int32),
                ("c_mults", C.c_int32 * 8), ("strides", C.c_int32 * 8), ("max_batch", C.c_int32),
                ("max_latent_len", C.c_int32), ("precision", C.c_int32), ("with_encoder", C.c_int32), ("in_channels", C.c_int32),
                ("enc_latent_dim", C.c_int32)]


class T5Desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vocab_size", "d_model", "d_kv", "num_heads", "d_ff", "num_layers", "num_buckets", "max_distance")] + \
               [("eps", C.c_float)] + [(n, C.c_int32) for n in ("max_batch", "max_len", "precision")]


class TestEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("bias_mod", C.c_int32), ("resid", C.c_void_p), ("ldr", C.c_int32),
                ("gate", C.c_void_p), ("gate_bstride", C.c_int32), ("rows_per_batch", C.c_int32),
                ("out_f32", C.c_void_p), ("ld32", C.c_int32), ("out_bf16", C.c_void_p), ("ld16", C.c_int32),
                ("split_stride", C.c_int32), ("act", C.c_int32), ("act_a", C.c_void_p), ("act_b", C.c_void_p)]


_lib = None

_VP, _I, _F = C.c_void_p, C.c_int, C.c_float
_SIGS = {
    "ezb_version": ([], C.c_int),
    "ezb_last_error": ([], C.c_char_p),
    "ezb_dit_create": ([C.POINTER(_VP), C.POINTER(DitDesc), _I], _I),
    "ezb_dit_destroy": ([_VP], _I),
    "ezb_dit_load_weight": ([_VP, C.c_char_p, _VP, C.POINTER(C.c_int64), _I, _VP], _I),
    "ezb_dit_finalize_weights": ([_VP, _VP], _I),
    "ezb_dit_set_context": ([_VP, _VP, _VP, _I, _I, _VP], _I),
    "ezb_dit_set_timesteps": ([_VP, C.POINTER(C.c_int64), _I, _VP], _I),
    "ezb_dit_forward": ([_VP, _VP, _VP, _VP, C.POINTER(C.c_int32), _I, C.POINTER(_VP), _VP, _I, _I, _VP], _I),
    "ezb_controlnet_forward": ([_VP, _VP, _VP, _VP, C.POINTER(C.c_int32), _I, _VP, _F, C.POINTER(_VP), _I, _I, _VP], _I),
    "ezb_cfg_ddim_step": ([_I, _VP, _VP, _VP, _I, _I, _I, _F, _F, C.POINTER(C.c_float), _VP], _I),
    "ezb_option_epoch": ([], C.c_ulonglong),
    "ezb_vae_create": ([C.POINTER(_VP), C.POINTER(VaeDesc), _I], _I),
    "ezb_vae_destroy": ([_VP], _I),
    "ezb_vae_load_weight": ([_VP, C.c_char_p, _VP, C.POINTER(C.c_int64), _I, _VP], _I),
    "ezb_vae_finalize_weights": ([_VP, _VP], _I),
    "ezb_vae_decode": ([_VP, _VP, _VP, _I, _I, _VP], _I),
    "ezb_vae_encode": ([_VP, _VP, _VP, _VP, _I, _I, _VP], _I),
    "ezb_t5_create": ([C.POINTER(_VP), C.POINTER(T5Desc), _I], _I),
    "ezb_t5_destroy": ([_VP], _I),
    "ezb_t5_load_weight": ([_VP, C.c_char_p, _VP, C.POINTER(C.c_int64), _I, _VP], _I),
    "ezb_t5_finalize_weights": ([_VP, _VP], _I),
    "ezb_t5_forward": ([_VP, _VP, _VP, _VP, _VP, _I, _I, _VP], _I),
    "ezb_energy_condition": ([_I, _VP, _VP, _I, _I, _I, _I, _F, _I, _I, _VP], _I),
    "ezb_wave_prepare": ([_I, _VP, _VP, _I, _I, _I, _I, _F, _VP], _I),
    "ezb_wave_splice": ([_I, _VP, C.c_longlong, _VP, C.c_longlong, C.c_longlong, _VP], _I),
    "ezb_wave_to_pcm16": ([_I, _VP, _VP, C.c_longlong, _VP], _I),
    "ezb_set_option": ([C.c_char_p, _I], _I),
    "ezb_debug_read": ([C.POINTER(C.c_ulonglong)], _I),
    "ezb_launch_count": ([], C.c_ulonglong),
    "ezb_launch_count_add": ([C.c_ulonglong], None),
    "ezb_prof_gemm_begin": ([], _I),
    "ezb_prof_gemm_end": ([C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)], _I),
    "ezb_prof_gemm_stats": ([C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)], _I),
    "ezb_test_gemm": ([_I, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, C.POINTER(TestEpilogue), _I, _I, _I, _I, _I, _I, _VP], _I),
    "ezb_test_attention": ([_I, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP], _I),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Loads the library (once).  Raises EzbError when it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EzbError(f"{LIB_PATH} not found: run `python -m ezaudio_b200.build` (or __graft_entry__.build()); "
                           "ezaudio_b200 has no CPU / PyTorch fallback")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.argtypes, fn.restype = args, res
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise EzbError(f"libezb200 error {rc}: {lib().ezb_last_error().decode()}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
