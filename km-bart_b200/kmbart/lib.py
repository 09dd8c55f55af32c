"""ctypes binding of libkmbart_sm100.so (C-ABI declared in include/kmbart.h).

The product path has NO fallback: if the shared library is missing or the device is not
sm_100, importing the kernels raises.  Build with `python __graft_entry__.py` / `make`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KMBART_LIB_PATH: an alternative build of the same library (A/B timing of kernel variants on one box)
LIB_PATH = os.environ.get("KMBART_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "libkmbart_sm100.so")

c_void_p, c_int, c_int64, c_float, c_uint32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32

EPI_LINEAR, EPI_CE_STATS, EPI_CE_GRAD = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_GELU_GRAD, ACT_TANH, ACT_TANH_GRAD = 0, 1, 2, 3, 4


class GemmEpilogue(C.Structure):
    """Mirror of struct KmbGemmEpilogue."""
    _fields_ = [
        ("mode", C.c_int32), ("act", C.c_int32), ("alpha", c_float), ("accumulate", C.c_int32),
        ("bias", c_void_p), ("residual", c_void_p), ("ld_res", c_int64),
        ("aux", c_void_p), ("ld_aux", c_int64),
        ("out_f32", c_void_p), ("ld_f32", c_int64),
        ("out_bf16", c_void_p), ("ld_bf16", c_int64), ("out_preact", c_void_p),
        ("dropout_p", c_float), ("dropout_tag", c_uint32), ("dropout_seed", c_void_p),
        ("labels", c_void_p), ("ce_max", c_void_p), ("ce_sum", c_void_p), ("ce_label_logit", c_void_p),
        ("ce_lse", c_void_p), ("ce_gscale", c_void_p),
    ]


DECODE_MAX_LAYERS = 12


class DecodeLayer(C.Structure):
    """Mirror of struct KmbDecodeLayer."""
    _fields_ = [(n, c_void_p) for n in (
        "w_qkv", "w_o", "w_cq", "w_co", "w_fc1", "w_fc2", "b_qkv", "b_o", "b_cq", "b_co", "b_fc1", "b_fc2",
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "ln3_g", "ln3_b", "cache", "cross_kv")] + [("packed", c_void_p * 6)]


class DecodeStep(C.Structure):
    """Mirror of struct KmbDecodeStep (persistent decode step, csrc/decode_mega.cu)."""
    _fields_ = ([(n, C.c_int32) for n in ("rows", "d", "H", "F", "L", "t", "max_len", "Se", "row_div", "pos_row")]
                + [("embed_scale", c_float), ("attn_scale", c_float), ("nt", C.c_int32 * 6)]
                + [(n, c_void_p) for n in ("ids", "tok_emb", "pos_emb", "lne_g", "lne_b", "slot_tbl", "key_pad",
                                           "x_f32", "x_b16", "ctx", "lin", "q2", "h", "barrier", "trace")]
                + [("layers", DecodeLayer * DECODE_MAX_LAYERS)])


class DecodeStepC(C.Structure):
    """Mirror of struct KmbDecodeStepC (cluster variant of the persistent decode step, csrc/decode_cluster.cu)."""
    _fields_ = ([(n, C.c_int32) for n in ("rows", "d", "H", "F", "L", "t", "max_len", "Se", "row_div", "pos_row")]
                + [("embed_scale", c_float), ("attn_scale", c_float), ("flags", C.c_int32), ("reserved", C.c_int32)]
                + [(n, c_void_p) for n in ("ids", "tok_emb", "pos_emb", "lne_g", "lne_b", "slot_tbl", "key_pad",
                                           "y0", "y1", "stats", "x_f32", "x_b16", "ctx", "h", "barrier", "trace")]
                + [("layers", DecodeLayer * DECODE_MAX_LAYERS)])


class BeamState(C.Structure):
    """Mirror of struct KmbBeamState."""
    _fields_ = ([(n, C.c_int32) for n in ("batch", "num_beams", "K", "V", "eos", "pad", "max_len", "early_stopping")]
                + [("length_penalty", C.c_double)]
                + [(n, c_void_p) for n in ("cand_val", "cand_tok", "beam_scores", "hist", "slot_tbl", "ids_next", "beam_idx", "done",
                                           "hyp_n", "hyp_score", "hyp_len", "hyp_tok", "worst", "done_count")])


# name -> argtypes; every function returns int except those listed in _RESTYPES
_PROTOS = {
    "kmb_version": [],
    "kmb_arch_check": [],
    "kmb_gemm": [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_int, c_int,
                 C.POINTER(GemmEpilogue), c_int, c_void_p],
    "kmb_gemm_n_tiles": [c_int, c_int],
    "kmb_gemm_pick_tile_n": [c_int, c_int],
    "kmb_gemm_debug_timeline": [c_int, c_int, c_void_p],
    "kmb_attn_fwd": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                     c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "kmb_attn_fwd_strided": [c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(c_int64), c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "kmb_decode_attn": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64,
                        c_void_p, c_int64, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "kmb_attn_f32": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64,
                     c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "kmb_decode_step": [C.POINTER(DecodeStep), c_void_p],
    "kmb_decode_step_grid": [],
    "kmb_decode_step_cluster": [C.POINTER(DecodeStepC), c_void_p],
    "kmb_decode_cluster_grid": [c_int],
    "kmb_decode_cluster_barriers": [c_int],
    "kmb_decode_pack_offsets": [c_int, c_int, c_int, C.POINTER(c_int64)],
    "kmb_decode_pack_weights": [C.POINTER(DecodeLayer), c_int, c_int, c_int, c_void_p, c_void_p],
    "kmb_greedy_select": [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                          c_void_p, c_void_p],
    "kmb_select_max_vocab": [],
    "kmb_sample_select": [c_void_p, c_int64, c_int, c_int, c_float, c_int, c_float, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_int64, c_void_p, c_void_p],
    "kmb_beam_step": [c_void_p, c_int64, C.POINTER(BeamState), c_int, c_int, c_int, c_void_p],
    "kmb_attn_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                     c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "kmb_pack_features": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "kmb_slot_index": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "kmb_embed_ln_fwd": [c_void_p] * 15 + [c_int, c_int, c_int, c_int, c_void_p, c_float, c_float, c_uint32,
                                           c_void_p, c_void_p],
    "kmb_embed_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                      c_float, c_int, c_void_p],
    "kmb_box_wgrad": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "kmb_layernorm_fwd": [c_void_p] * 9 + [c_int, c_int, c_float, c_uint32, c_void_p, c_void_p],
    "kmb_layernorm_bwd": [c_void_p] * 11 + [c_int, c_int, c_float, c_uint32, c_float, c_uint32, c_void_p, c_void_p],
    "kmb_colsum_bf16": [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p],
    "kmb_gather_rows_bf16": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p],
    "kmb_scatter_add_rows": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p],
    "kmb_ce_combine": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float,
                       c_void_p, c_void_p, c_int, c_void_p],
    "kmb_ce_gscale": [c_void_p, c_void_p, c_float, c_void_p, c_void_p],
    "kmb_small_xent": [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_float, c_void_p,
                       c_void_p, c_int64, c_void_p, c_void_p],
    "kmb_adamw_chunk_elems": [],
    "kmb_adamw_multi": [c_void_p, c_void_p, c_int, c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_int,
                        c_void_p, c_void_p],
    "kmb_adamw_multi_part": [c_void_p, c_void_p, c_int, c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_int,
                             c_void_p, c_int, c_void_p],
    "kmb_ipc_export": [c_void_p, c_void_p, C.POINTER(C.c_ulonglong)],
    "kmb_ipc_open": [c_void_p, C.POINTER(c_void_p)],
    "kmb_ipc_close": [c_void_p],
    "kmb_peer_can_access": [c_int, c_int],
    "kmb_peer_ctx_create": [c_int, c_int, c_void_p, C.POINTER(c_void_p), c_void_p, C.POINTER(c_void_p), c_void_p, C.POINTER(c_void_p),
                            C.c_size_t, c_int, C.POINTER(c_void_p)],
    "kmb_peer_ctx_destroy": [c_void_p],
    "kmb_peer_exchange_region": [c_void_p, c_int, C.c_size_t, C.c_size_t, c_uint32, c_int, c_void_p],
    "kmb_peer_join": [c_void_p, c_void_p],
    "kmb_peer_mark": [c_void_p],
    "kmb_peer_join_mark": [c_void_p, c_void_p],
    "kmb_cast_bf16": [c_void_p, c_void_p, c_int64, c_void_p],
    "kmb_repack_img_weight": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "kmb_invert_mask": [c_void_p, c_void_p, c_int64, c_void_p],
    "kmb_next_seed": [c_void_p, c_void_p, c_void_p],
    "kmb_split_tf32": [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p],
}
_RESTYPES = {"kmb_last_error": C.c_char_p}

EXPORTED_SYMBOLS = sorted(list(_PROTOS) + ["kmb_last_error"])

_lib = None


class KmbartError(RuntimeError):
    pass


def load():
    """Load the library once; raise (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KmbartError(
            f"{LIB_PATH} not found: the sm_100a CUDA library is required (no CPU fallback). "
            "Build it with `python __graft_entry__.py` or `make` at the repo root.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    lib.kmb_last_error.argtypes = []
    lib.kmb_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().kmb_last_error().decode(errors="replace")
        raise KmbartError(f"{what} failed with code {rc}: {msg}")


_arch_ok = False


def require_b200():
    """Hard gate used by every GPU entry point."""
    global _arch_ok
    if not _arch_ok:
        check(load().kmb_arch_check(), "kmb_arch_check")
        _arch_ok = True
