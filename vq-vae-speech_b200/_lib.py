"""ctypes binding of libvqs_b200.so (the C ABI declared in include/vqs_b200.h).

The product path has NO fallback: if the shared library is missing and cannot be built, importing an op raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libvqs_b200.so')

LAYOUT_FLAT_ND = 0
LAYOUT_BDT_AS_DTB = 1
PREC_FP32, PREC_TF32, PREC_TF32X3 = 0, 1, 2
PRECISIONS = {'fp32': PREC_FP32, 'tf32': PREC_TF32, '3xtf32': PREC_TF32X3}


class ConvGemmDesc(Structure):
    """struct vqs_conv_gemm_desc (include/vqs_b200.h)."""
    _fields_ = [
        ('A', c_void_p), ('X', c_void_p), ('a_tap_major', c_int),
        ('M', c_int), ('Cred', c_int), ('ksz', c_int),
        ('B', c_int), ('Lin', c_int), ('Lout', c_int),
        ('x_sb', c_longlong), ('x_sc', c_longlong), ('x_sl', c_longlong),
        ('l_mul', c_int), ('j_mul', c_int), ('off', c_int), ('l_div', c_int),
        ('x_relu', c_int),
        ('bias', c_void_p), ('add_pre', c_void_p), ('add_pre_relu', c_int), ('relu', c_int),
        ('mask_out', c_void_p), ('mask', c_void_p), ('mask_kind', c_int),
        ('add_post', c_void_p), ('out', c_void_p), ('out2', c_void_p), ('mask2', c_void_p), ('mask2_kind', c_int),
        ('precision', c_int),
        ('splitk_ws', c_void_p), ('splitk_ws_bytes', c_size_t),
    ]


class WgradDesc(Structure):
    """struct vqs_wgrad_desc (include/vqs_b200.h)."""
    _fields_ = [
        ('Aact', c_void_p), ('X', c_void_p),
        ('M', c_int), ('Cred', c_int), ('ksz', c_int),
        ('B', c_int), ('La', c_int), ('Lx', c_int),
        ('l_mul', c_int), ('j_mul', c_int), ('off', c_int),
        ('x_relu', c_int),
        ('dW', c_void_p), ('accumulate', c_int), ('precision', c_int),
    ]


class PermuteItem(Structure):
    """struct vqs_permute_item (include/vqs_b200.h)."""
    _fields_ = [('w', c_void_p), ('out', c_void_p), ('d0', c_int), ('d1', c_int), ('k', c_int), ('mode', c_int)]


PERMUTE_MAX_ITEMS = 32

DP_MAX_WORLD, DP_CHANNELS, DP_PAD_WORD0, DP_EPOCH_WORDS = 8, 4, 256, 8


class DpCtx(Structure):
    """struct vqs_dp_ctx (include/vqs_b200.h)."""
    _fields_ = [('rank', c_int), ('world', c_int), ('peer_pads', c_void_p * DP_MAX_WORLD), ('epochs', c_void_p)]


class DpPtrs(Structure):
    """struct vqs_dp_ptrs (include/vqs_b200.h)."""
    _fields_ = [('p', c_void_p * DP_MAX_WORLD)]

# name -> (restype, argtypes); every symbol include/vqs_b200.h declares
PROTOTYPES = {
    'vqs_version': (c_int, []),
    'vqs_last_error': (c_char_p, []),
    'vqs_launch_count': (c_longlong, []),
    'vqs_engine_count': (c_longlong, [c_int]),
    'vqs_vq_workspace_bytes': (c_size_t, [c_int, c_int]),
    'vqs_vq_assign': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_size_t, c_void_p]),
    'vqs_vq_set_engine': (c_int, [c_int]),
    'vqs_vq_one_hot': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    'vqs_vq_ema_update': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int,
                                  c_int, c_void_p]),
    'vqs_vq_quantize': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                c_void_p, c_double, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    'vqs_vq_backward': (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_int, c_void_p, c_void_p]),
    'vqs_vq_backward_loss': (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_int, c_void_p, c_void_p, c_double, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    'vqs_vq_grad_codebook': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_int, c_void_p]),
    'vqs_conv_gemm': (c_int, [POINTER(ConvGemmDesc), c_void_p]),
    'vqs_wgrad_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'vqs_wgrad_gemm': (c_int, [POINTER(WgradDesc), c_void_p, c_size_t, c_void_p]),
    'vqs_bias_grad': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'vqs_permute_weight': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'vqs_permute_weights': (c_int, [POINTER(PermuteItem), c_int, c_void_p]),
    'vqs_upsample2_fwd': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    'vqs_upsample2_bwd': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    'vqs_jitter_fwd': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    'vqs_jitter_bwd': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    'vqs_mse_fwd_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_longlong, c_longlong, c_float,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'vqs_relu_fwd': (c_int, [c_void_p, c_longlong, c_void_p, c_void_p]),
    'vqs_relu_bwd': (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    'vqs_add': (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    'vqs_scale': (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    'vqs_blc_to_ncl': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'vqs_concat_channels': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'vqs_slice_channels': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'vqs_pairwise_l2': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'vqs_weight_norm_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'vqs_weight_norm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'vqs_normalize_features': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    'vqs_dp_barrier': (c_int, [c_void_p, c_int, c_void_p]),
    'vqs_dp_allreduce_small': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    'vqs_dp_amsgrad_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                    c_void_p, c_int, c_double, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    'vqs_dp_amsgrad_range': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                     c_longlong, c_void_p, c_int, c_double, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    'vqs_amsgrad_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_int,
                                 c_double, c_double, c_double, c_double, c_double, c_void_p]),
}

_LIB = None


def load():
    """Loads (building first if the sources changed and nvcc is present) libvqs_b200.so.  Raises when unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get('VQS_LIB_PATH')      # profiling builds (profiles/build_variant.py); default: the in-tree library
    if not path:
        from . import build as _build
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError('libvqs_b200.so is missing (%s); run `python vq-vae-speech_b200/build.py`' % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(code):
    if code != 0:
        msg = load().vqs_last_error()
        raise RuntimeError('libvqs_b200 error %d: %s' % (code, msg.decode() if msg else '?'))


def launch_count():
    return int(load().vqs_launch_count())


ENGINES = {'conv_cudacore': 0, 'conv_tc': 1, 'wgrad_cudacore': 2, 'wgrad_tc': 3, 'wgrad_tma': 4}


def engine_counts():
    """{engine name: GEMM calls dispatched to it so far in this process} (include/vqs_b200.h: VQS_ENGINE_*)."""
    lib = load()
    return dict((k, int(lib.vqs_engine_count(v))) for k, v in ENGINES.items())
