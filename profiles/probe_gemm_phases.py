"""Phase stamps of the tcgen05 GEMM kernel (profiling build only):
    VQS_EXTRA_NVCC_FLAGS=-DVQS_GEMM_TIMING python vq-vae-speech_b200/build.py --force
    VQS_EXTRA_NVCC_FLAGS=-DVQS_GEMM_TIMING python profiles/probe_gemm_phases.py
CTA (0, 0, 0) of the last launch: cycles from kernel entry to the end of the prologue (barriers, TMEM, griddepcontrol.wait), to
the first full stage, through the MMA issue loop, to complete accumulators, through the epilogue."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import functional as F, ops, _lib  # noqa: E402

lib = _lib.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)


def stamps():
    buf = (ctypes.c_longlong * 48)()
    rc = lib.vqs_debug_gemm_timing(buf)
    assert rc == 0
    return list(buf)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(tag, ms, nkb):
    t = stamps()
    ns = t[7] - t[6]
    cyc = t[4] - t[5]
    print('%-44s %.4f ms | CTA0: %6d ns = %7d cyc (%.0f MHz) | prologue %6d | to first full %6d | issue loop %7d (%5.0f / k-block) | '
          'drain %6d | epilogue %6d' % (tag, ms, ns, cyc, cyc / max(ns, 1) * 1e3, t[0] - t[5], t[1] - t[0], t[2] - t[1],
                                        (t[2] - t[1]) / nkb, t[3] - t[2], t[4] - t[3]), flush=True)
    print('      epilogue of warp 0: TMEM loads + sums %d | smem transpose %d | row batches %d %d %d %d' % (
        t[8] - t[3], t[9] - t[8], t[10] - t[9], t[11] - t[10], t[12] - t[11], t[13] - t[12]), flush=True)
    print('      issuers waiting for a full stage: main %d, corrections %d | producer thread 0 (prologue done -> last stage %d): '
          'waiting for an empty stage %d, stores (incl. waiting for the loads) %d, fence + arrive %d' % (
              t[14], t[15], t[19] - t[0], t[16], t[17], t[18]), flush=True)
    if nkb > 45:
        z = t[20]
        print('      stage 0, k-block 40 (cycles after its slot was seen empty): producers arrived %d | issuers saw it full: main %d, corr %d | '
              'commits issued: main %d, corr %d | k-block 41 seen full: main %d corr %d | slot seen empty again (k-block 44) %d' % (
                  t[21] - z, t[23] - z, t[25] - z, t[24] - z, t[26] - z, t[27] - z, t[28] - z, t[22] - z), flush=True)
        print('      peer CTA, same k-block (cycles after ITS slot was seen empty): producers arrived %d | relay warp saw the stage full %d' % (
            t[30] - t[29], t[31] - t[29]), flush=True)
        print('      leader CTA, the four warps of the group: slot seen empty %s | stores done %s | arrived %s' % (
            [t[36 + w] - z for w in range(4)], [t[40 + w] - z for w in range(4)], [t[32 + w] - z for w in range(4)]), flush=True)


ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else '3xtf32')
for (B, C, L, k) in ((64, 768, 47, 3), (64, 768, 48, 3), (64, 768, 24, 3), (64, 768, 48, 1)):
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b = torch.randn(C, device=dev)
    res = torch.randn(B, C, L, device=dev)
    mask_out = torch.empty(B, C, L, dtype=torch.uint8, device=dev)
    y = torch.empty(B, C, L, device=dev)
    dW = torch.empty_like(w)
    ws = F._wgrad_ws(C, C, k, B, L, dev)
    A = F.gemm_weight(w, 'conv_fwd')
    pad = k // 2
    tag = 'B%d C%d L%d k%d' % (B, C, L, k)
    nkb = C * k // 32
    ms = timeit(lambda: F.conv1d_forward(x, A, b, 1, pad, out=y))
    report(tag + ' conv fwd', ms, nkb if not (L == 24 and k == 3) else nkb / 2)
    ms = timeit(lambda: F.conv1d_forward(x, A, b, 1, pad, out=y, relu=True, mask_out=mask_out, add_post=res))
    report(tag + ' conv fwd + relu/mask/skip', ms, nkb if not (L == 24 and k == 3) else nkb / 2)
    if k == 3:
        ms = timeit(lambda: F.conv1d_wgrad(y, x, dW, 1, pad, ws))
        t = stamps()
        print('%-44s %.4f ms | CTA0: to first full %d | issue loop %7d cyc | drain %d | epilogue %6d | total %7d | issuers waiting: main %d corr %d | '
              'producer thread 0: waiting for an empty stage %d, stores %d, fence + arrive %d' % (
                  tag + ' wgrad', ms, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[4] - t[0], t[14], t[15], t[16], t[17], t[18]), flush=True)
