"""CPU checks of the drop-in boundary: libvqs_b200.so loads without a GPU and exports every symbol that
include/vqs_b200.h declares (no compute calls), the ctypes structures match the C structs, argument validation works."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'vqs_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vqs_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ('vqs_vq_assign', 'vqs_vq_ema_update', 'vqs_vq_quantize', 'vqs_vq_backward', 'vqs_vq_grad_codebook',
                 'vqs_conv_gemm', 'vqs_wgrad_gemm', 'vqs_jitter_fwd', 'vqs_mse_fwd_bwd', 'vqs_amsgrad_step'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from vq_vae_speech_b200 import _lib
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), 'libvqs_b200.so does not export %s' % name
        assert name in _lib.PROTOTYPES, 'no ctypes prototype for %s' % name
    assert lib.vqs_version() >= 100


def test_struct_layout_matches_c():
    """sizeof / offsetof of the two descriptor structs as the C compiler sees them == the ctypes mirror."""
    from vq_vae_speech_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "vqs_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(vqs_conv_gemm_desc), offsetof(vqs_conv_gemm_desc, x_sb),
         offsetof(vqs_conv_gemm_desc, out), offsetof(vqs_conv_gemm_desc, precision),
         sizeof(vqs_wgrad_desc), offsetof(vqs_wgrad_desc, dW), offsetof(vqs_conv_gemm_desc, splitk_ws),
         offsetof(vqs_conv_gemm_desc, splitk_ws_bytes), sizeof(vqs_permute_item),
         sizeof(vqs_dp_ctx), offsetof(vqs_dp_ctx, peer_pads), offsetof(vqs_dp_ctx, epochs), sizeof(vqs_dp_ptrs));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 't.c')
        open(c, 'w').write(prog)
        exe = os.path.join(d, 't')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    C, W = _lib.ConvGemmDesc, _lib.WgradDesc
    assert vals == [ctypes.sizeof(C), C.x_sb.offset, C.out.offset, C.precision.offset, ctypes.sizeof(W), W.dW.offset,
                    C.splitk_ws.offset, C.splitk_ws_bytes.offset, ctypes.sizeof(_lib.PermuteItem),
                    ctypes.sizeof(_lib.DpCtx), _lib.DpCtx.peer_pads.offset, _lib.DpCtx.epochs.offset,
                    ctypes.sizeof(_lib.DpPtrs)]
    assert (_lib.DP_MAX_WORLD, _lib.DP_CHANNELS, _lib.DP_PAD_WORD0) == (8, 4, 256)      # VQS_DP_* of the header


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device (returns VQS_ERR_ARG + a message)."""
    from vq_vae_speech_b200 import _lib
    lib = _lib.load()
    rc = lib.vqs_vq_one_hot(None, 0, 0, None, None)
    assert rc == 10001
    assert b'vqs_vq_one_hot' in lib.vqs_last_error()
    assert lib.vqs_vq_workspace_bytes(44, 64) > 0 and lib.vqs_vq_workspace_bytes(0, 64) == 0
    assert lib.vqs_wgrad_workspace_bytes(768, 768, 3, 64, 47) > 0
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_ops_refuse_cpu_tensors():
    import torch
    from vq_vae_speech_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU path'):
        ops.relu_fwd(torch.zeros(4))
