"""GPU parity tests of the conv-like implicit-GEMM kernels (forward / dgrad / wgrad of nn.Conv1d and nn.ConvTranspose1d,
fused epilogues) and the element-wise pieces, against the numpy oracle (oracle/model_oracle.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-5
# GEMM engines: exact-fp32 CUDA cores, tcgen05 3xTF32 (same 1e-5 bar), tcgen05 single-pass TF32 (stated tolerance: the
# 10-bit TF32 mantissa gives ~5e-4 relative error per product; 3e-3 of the output scale bounds the accumulated error)
PRECS = [('fp32', 1e-5), ('3xtf32', 1e-5), ('tf32', 3e-3)]


@pytest.fixture(params=PRECS, ids=[p[0] for p in PRECS])
def prec(request):
    from vq_vae_speech_b200 import ops
    prev = ops.set_precision(request.param[0])
    yield request.param
    ops.set_precision(prev)


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def _t(a, dev, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype))).to(dev)


CONV_CASES = [  # (B, Cin, Cout, L, k, stride, pad)
    (2, 39, 48, 47, 3, 1, 1), (2, 48, 48, 47, 4, 2, 2), (3, 48, 64, 24, 3, 1, 1), (2, 96, 80, 24, 1, 1, 0),
    (5, 64, 768, 24, 3, 1, 1), (64, 768, 768, 24, 3, 1, 1), (16, 768, 64, 96, 3, 1, 1), (2, 7, 2, 9, 3, 1, 1),
    (4, 130, 200, 47, 4, 2, 2), (1, 768, 768, 5, 3, 1, 1), (3, 64, 96, 47, 4, 2, 2), (9, 256, 160, 31, 2, 1, 0),
    # whole 128-row tiles, unit stride, L % 4 == 0, k = 1: shapes the TMA-fed wgrad kernel (wgrad_tma.cu) accepts
    (4, 128, 256, 48, 1, 1, 0), (3, 256, 128, 20, 1, 1, 0), (70, 128, 128, 36, 1, 1, 0), (64, 768, 768, 24, 1, 1, 0),
    (4, 128, 256, 48, 3, 1, 1),
    # channel counts that are no multiple of the 32-wide k-block under >= 128 output rows: tcgen05 through operand images padded
    # with zero channels (the 39 MFCC channels of the 768-wide model; dgrad of the second case: M = 65 stays on CUDA cores)
    (8, 39, 768, 47, 3, 1, 1), (3, 65, 256, 33, 2, 1, 0),
]


@pytest.mark.parametrize('B,Cin,Cout,L,k,stride,pad', CONV_CASES)
def test_conv1d_fwd_dgrad_wgrad(B, Cin, Cout, L, k, stride, pad, prec):
    dev = _dev()
    TOL = prec[1]
    from vq_vae_speech_b200 import functional as F, ops
    rng = np.random.RandomState(B * 31 + Cin)
    x = rng.randn(B, Cin, L)
    w = rng.randn(Cout, Cin, k) / np.sqrt(Cin * k)
    b = rng.randn(Cout)
    y_o = mo.conv1d_fwd(x, w, b, stride, pad)
    xd, wd, bd = _t(x, dev), _t(w, dev), _t(b, dev)
    y = F.conv1d_forward(xd, F.gemm_weight(wd, 'conv_fwd'), bd, stride, pad)   # operand image when Cin % 32 == 0
    assert tuple(y.shape) == y_o.shape
    assert rel_err(y.cpu().numpy(), y_o) < TOL
    gy = rng.randn(*y_o.shape)
    gyd = _t(gy, dev)
    dx = F.conv1d_dgrad(gyd, F.gemm_weight(wd, 'conv_dgrad'), L, stride, pad)
    assert rel_err(dx.cpu().numpy(), mo.conv1d_dgrad(gy, w, L, stride, pad)) < TOL
    dW = torch.empty_like(wd)
    F.conv1d_wgrad(gyd, xd, dW, stride, pad, F._wgrad_ws(Cout, Cin, k, B, y.shape[2], dev))
    dw_o, db_o = mo.conv1d_wgrad(gy, x, k, stride, pad)
    assert rel_err(dW.cpu().numpy(), dw_o) < TOL
    db = ops.bias_grad(gyd, torch.empty(Cout, device=dev))
    assert rel_err(db.cpu().numpy(), db_o) < TOL
    # accumulate flag
    F.conv1d_wgrad(gyd, xd, dW, stride, pad, F._wgrad_ws(Cout, Cin, k, B, y.shape[2], dev), accumulate=True)
    assert rel_err(dW.cpu().numpy(), 2 * dw_o) < TOL


CONVT_CASES = [  # (B, Cin, Cout, L, k, pad)
    (2, 48, 48, 48, 3, 1), (2, 48, 48, 48, 3, 0), (2, 48, 39, 50, 2, 0), (16, 768, 39, 50, 2, 0), (8, 768, 768, 48, 3, 0),
]


@pytest.mark.parametrize('B,Cin,Cout,L,k,pad', CONVT_CASES)
def test_conv_transpose1d_fwd_dgrad_wgrad(B, Cin, Cout, L, k, pad, prec):
    dev = _dev()
    TOL = prec[1]
    from vq_vae_speech_b200 import functional as F, ops
    rng = np.random.RandomState(B * 17 + Cout)
    x = rng.randn(B, Cin, L)
    w = rng.randn(Cin, Cout, k) / np.sqrt(Cin * k)
    b = rng.randn(Cout)
    y_o = mo.convT1d_fwd(x, w, b, pad)
    xd, wd, bd = _t(x, dev), _t(w, dev), _t(b, dev)
    A = F.gemm_weight(wd, 'convT_fwd')
    y = F.convT1d_forward(xd, A, bd, pad)
    assert tuple(y.shape) == y_o.shape
    assert rel_err(y.cpu().numpy(), y_o) < TOL
    # trimmed output (convolutional_vq_vae.py:133-137 drops the tail): only the first `keep` positions are computed,
    # and the backward sees zero gradient beyond them
    keep = y_o.shape[2] - 3
    yt = F.convT1d_forward(xd, A, bd, pad, out_len=keep)
    assert rel_err(yt.cpu().numpy(), y_o[:, :, :keep]) < TOL
    gy = rng.randn(B, Cout, keep)
    gfull = np.zeros_like(y_o)
    gfull[:, :, :keep] = gy
    gyd = _t(gy, dev)
    dx = F.convT1d_dgrad(gyd, F.gemm_weight(wd, 'convT_dgrad'), L, pad)
    assert rel_err(dx.cpu().numpy(), mo.convT1d_dgrad(gfull, w, L, pad)) < TOL
    dW = torch.empty_like(wd)
    F.convT1d_wgrad(gyd, xd, dW, pad, F._wgrad_ws(Cin, Cout, k, B, L, dev))
    dw_o, db_o = mo.convT1d_wgrad(gfull, x, k, pad)
    assert rel_err(dW.cpu().numpy(), dw_o) < TOL
    assert rel_err(ops.bias_grad(gyd, torch.empty(Cout, device=dev)).cpu().numpy(), db_o) < TOL


def test_fused_epilogue_and_strided_input(prec):
    dev = _dev()
    TOL = prec[1]
    from vq_vae_speech_b200 import functional as F, ops
    rng = np.random.RandomState(0)
    B, Cin, Cout, L = 3, 39, 64, 47
    x_blc = rng.randn(B, L, Cin)                      # (B, T, F) feature batch, read through strides (no permute copy)
    w = rng.randn(Cout, Cin, 3) / 10
    b = rng.randn(Cout)
    res = rng.randn(B, Cout, L)
    pre = rng.randn(B, Cout, L)
    xd = _t(x_blc, dev)
    mask_out = torch.empty(B, Cout, L, dtype=torch.uint8, device=dev)
    y = F.conv1d_forward(xd, F.gemm_weight(_t(w, dev), 'conv_fwd'), _t(b, dev), 1, 1, x_strides=(L * Cin, 1, Cin),
                         x_shape=(B, Cin, L),
                         add_pre=_t(pre, dev), add_pre_relu=True, relu=True, mask_out=mask_out, add_post=_t(res, dev))
    p = mo.conv1d_fwd(x_blc.transpose(0, 2, 1), w, b, 1, 1) + np.maximum(pre, 0)
    r = np.maximum(p, 0)
    assert rel_err(y.cpu().numpy(), r + res) < TOL
    m = mask_out.cpu().numpy().astype(bool)
    safe = np.abs(p) > 1e-4
    assert np.array_equal(m[safe], (p > 0)[safe])
    # x_relu + mask (float / uint8) + out2
    x = rng.randn(B, Cout, L)
    w2 = rng.randn(Cout, Cout, 3) / 10
    act = rng.randn(B, Cout, L)
    out2 = torch.empty(B, Cout, L, device=dev)
    y2 = F.conv1d_forward(_t(x, dev), F.gemm_weight(_t(w2, dev), 'conv_fwd'), None, 1, 1, x_relu=True, mask=_t(act, dev), mask_kind=ops.MASK_FLOAT,
                          add_post=_t(res, dev), out2=out2, mask2=mask_out, mask2_kind=ops.MASK_U8)
    v = mo.conv1d_fwd(np.maximum(x, 0), w2, None, 1, 1) * (act > 0) + res
    assert rel_err(y2.cpu().numpy(), v) < TOL
    assert rel_err(out2.cpu().numpy(), v * m) < TOL


def test_elementwise_ops():
    dev = _dev()
    from vq_vae_speech_b200 import ops
    rng = np.random.RandomState(1)
    x = rng.randn(3, 10, 24)
    xd = _t(x, dev)
    assert np.array_equal(ops.upsample2_fwd(xd).cpu().numpy(), mo.upsample2(x).astype(np.float32))
    g = rng.randn(3, 10, 48)
    assert rel_err(ops.upsample2_bwd(_t(g, dev)).cpu().numpy(), mo.upsample2_bwd(g)) < TOL
    np.random.seed(7)
    src = mo.jitter_plan(24, 0.5)
    srcd = torch.from_numpy(src.astype(np.int32)).to(dev)
    assert np.array_equal(ops.jitter_fwd(xd, srcd).cpu().numpy(), x[:, :, src].astype(np.float32))
    keep = (src == np.arange(24))
    assert np.array_equal(ops.jitter_bwd(xd, srcd).cpu().numpy(), (x * keep[None, None, :]).astype(np.float32))
    assert np.array_equal(ops.relu_fwd(xd).cpu().numpy(), np.maximum(x, 0).astype(np.float32))
    assert np.array_equal(ops.blc_to_ncl(xd).cpu().numpy(), x.transpose(0, 2, 1).astype(np.float32))
    w = rng.randn(5, 7, 3)
    assert np.array_equal(ops.permute_weight(_t(w, dev)).cpu().numpy(), w.transpose(1, 0, 2).astype(np.float32))
    assert np.array_equal(ops.permute_weight(_t(w, dev), mode=1).cpu().numpy(), w.transpose(0, 2, 1).astype(np.float32))
    assert np.array_equal(ops.permute_weight(_t(w, dev), mode=2).cpu().numpy(), w.transpose(1, 2, 0).astype(np.float32))
    # MSE forward + backward against a strided (B, T, F) target
    recon = rng.randn(3, 13, 24)
    tgt_btf = rng.randn(3, 24, 13)
    loss = torch.empty(1, device=dev)
    grad = torch.empty(3, 13, 24, device=dev)
    ops.mse_fwd_bwd(_t(recon, dev), _t(tgt_btf, dev), (24 * 13, 1, 13), 1.0, loss, grad, ops.mse_workspace(dev))
    diff = recon - tgt_btf.transpose(0, 2, 1)
    assert rel_err(loss.item(), np.mean(diff ** 2)) < TOL
    assert rel_err(grad.cpu().numpy(), 2 * diff / diff.size) < TOL


def test_amsgrad_matches_oracle_and_torch():
    dev = _dev()
    from vq_vae_speech_b200 import ops
    rng = np.random.RandomState(2)
    n = 100003
    p = rng.randn(n)
    po = p.copy()
    m = np.zeros(n); v = np.zeros(n); vm = np.zeros(n)
    pd, md, vd, vmd = _t(p, dev), _t(m, dev), _t(v, dev), _t(vm, dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for s in range(1, 4):
        g = rng.randn(n) * (10.0 ** rng.randint(-4, 1))
        po, m, v, vm = mo.amsgrad_step(po, g, m, v, vm, s, 2e-4)
        ops.amsgrad_step(pd, _t(g, dev), md, vd, vmd, step, 2e-4)
    assert int(step.item()) == 3
    # fp32 parameters vs the fp64 oracle: within 2 ulp of the largest parameter (the updates are ~lr = 2e-4 per step)
    assert float(np.max(np.abs(pd.cpu().numpy() - po))) < 2 * 1.2e-7 * np.abs(po).max() + 1e-5 * 2e-4 * 3
    assert rel_err(vmd.cpu().numpy(), vm) < TOL


def test_permute_weights_batched_matches_single_calls():
    """vqs_permute_weights (one launch for all the step's GEMM operands) == vqs_permute_weight item by item, every mode
    and kernel size, including operand images whose M is not a multiple of 128 (zero rows) and more than one chunk of
    VQS_PERMUTE_MAX_ITEMS items."""
    dev = _dev()
    from vq_vae_speech_b200 import ops
    rng = np.random.RandomState(3)
    items, want = [], []
    for rep in range(3):
        for (d0, d1, k) in [(5, 7, 3), (64, 96, 1), (39, 64, 2), (200, 32, 4), (128, 64, 3), (33, 65, 2)]:
            w = _t(rng.randn(d0, d1, k), dev)
            for mode in range(5):
                # (operand images of widths that are no multiple of 32 are padded with zero channels: 7 -> 32, 65 -> 96)
                ref = ops.permute_weight(w, mode=mode)
                out = torch.full_like(ref, float('nan'))
                items.append((w, out, mode))
                want.append(ref)
    assert len(items) > 64
    ops.permute_weights(items)
    torch.cuda.synchronize()
    for (w, out, mode), ref in zip(items, want):
        assert torch.equal(out, ref), 'mode %d shape %s' % (mode, tuple(w.shape))


@pytest.mark.parametrize('B,C,M,L', [(4, 128, 256, 48), (3, 256, 128, 20), (70, 128, 128, 36), (64, 768, 768, 24)])
def test_wgrad_tma_engine_matches_oracle(B, C, M, L, monkeypatch):
    """The TMA-fed weight-gradient kernel (default on; VQS_WGRAD_TMA=0 disables it: tensor-map boxes as raw tf32 operands, threads only
    derive the lo tiles) on the shapes it accepts (1 x 1 convolutions), incl. the fused input ReLU and accumulation."""
    dev = _dev()
    from vq_vae_speech_b200 import functional as F, ops
    monkeypatch.setenv('VQS_WGRAD_TMA', '1')
    prev = ops.set_precision('3xtf32')
    try:
        rng = np.random.RandomState(B + C)
        x, gy = rng.randn(B, C, L), rng.randn(B, M, L)
        xd, gyd = _t(x, dev), _t(gy, dev)
        dW = torch.empty(M, C, 1, device=dev)
        ws = F._wgrad_ws(M, C, 1, B, L, dev)
        F.conv1d_wgrad(gyd, xd, dW, 1, 0, ws)
        dw_o, _ = mo.conv1d_wgrad(gy, x, 1, 1, 0)
        assert rel_err(dW.cpu().numpy(), dw_o) < 1e-5
        F.conv1d_wgrad(gyd, xd, dW, 1, 0, ws, accumulate=True)
        assert rel_err(dW.cpu().numpy(), 2 * dw_o) < 1e-5
    finally:
        ops.set_precision(prev)


@pytest.mark.parametrize('B,Cin,Cout,L,k', [(64, 768, 64, 24, 3), (16, 256, 39, 48, 3), (8, 512, 128, 40, 1)])
def test_conv_forward_split_k_matches_oracle(B, Cin, Cout, L, k):
    """Few-tile conv GEMMs (M <= 128) with a bias-only epilogue and a scratch buffer split the reduction over several CTAs
    (gemm_tc.cu: partial tiles + conv_splitk_epilogue_kernel); same 1e-5 bar, and identical to the unsplit result's
    oracle."""
    dev = _dev()
    from vq_vae_speech_b200 import functional as F, ops
    prev = ops.set_precision('3xtf32')
    try:
        rng = np.random.RandomState(Cin + Cout)
        x = rng.randn(B, Cin, L)
        w = rng.randn(Cout, Cin, k) / np.sqrt(Cin * k)
        b = rng.randn(Cout)
        pad = (k - 1) // 2
        y_o = mo.conv1d_fwd(x, w, b, 1, pad)
        xd, wd, bd = _t(x, dev), _t(w, dev), _t(b, dev)
        ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        A = F.gemm_weight(wd, 'conv_fwd')
        n0 = ops._lib.launch_count()
        y = F.conv1d_forward(xd, A, bd, 1, pad, splitk_ws=ws)
        assert ops._lib.launch_count() - n0 == 2, 'expected the split kernel + its epilogue'
        assert rel_err(y.cpu().numpy(), y_o) < 1e-5
        y2 = F.conv1d_forward(xd, A, bd, 1, pad)                       # no scratch: single pass
        assert rel_err(y2.cpu().numpy(), y_o) < 1e-5
        y3 = F.conv1d_forward(xd, A, bd, 1, pad, relu=True, splitk_ws=ws)   # not a bias-only epilogue: no split
        assert rel_err(y3.cpu().numpy(), np.maximum(y_o, 0)) < 1e-5
    finally:
        ops.set_precision(prev)
