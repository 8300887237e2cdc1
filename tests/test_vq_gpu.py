"""GPU parity tests of the VQ bottleneck: CUDA path (through the C ABI) vs the CPU oracle and the golden vectors that the
unmodified reference produced (tests/golden/make_golden.py).  Run with `-m gpu` on a B200."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, rel_err
from oracle import vq_oracle as vqo

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north_star: losses, quantized outputs and gradients within 1e-5 relative (fp32)

VQ_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'vq_*.npz')))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _make_module(g, dev):
    from vq_vae_speech_b200.vector_quantizer import VectorQuantizer, VectorQuantizerEMA
    K, D = int(g['K']), int(g['D'])
    if bool(g['ema']):
        vq = VectorQuantizerEMA(K, D, float(g['commitment_cost']), float(g['decay']), dev, float(g['epsilon']))
    else:
        vq = VectorQuantizer(K, D, float(g['commitment_cost']), dev)
    vq = vq.to(dev)
    with torch.no_grad():
        vq._embedding.weight.copy_(_t(g['W0'], dev))
        if bool(g['ema']):
            vq._ema_w.copy_(_t(g['ema_w0'], dev))
            vq._ema_cluster_size.copy_(_t(g['cs0'], dev))
    return vq


@pytest.mark.parametrize('case', VQ_CASES)
def test_vq_module_matches_reference_golden(case):
    """Same inputs, same initial state -> the drop-in module reproduces what the reference module returned."""
    dev = _dev()
    g = load_golden(case)
    K, D, B, T = int(g['K']), int(g['D']), int(g['B']), int(g['T'])
    is_ema = bool(g['ema'])
    beta = float(g['commitment_cost'])
    vq = _make_module(g, dev).train()
    vq.record_near_ties = True
    W = g['W0']
    n_excluded = 0
    for s in range(int(g['steps'])):
        z = _t(g[f'z{s}'], dev).requires_grad_(True)
        gq = _t(g[f'g{s}'], dev)
        W_before = vq._embedding.weight.detach().cpu().numpy().copy()
        outs = vq(z, record_codebook_stats=True)
        vq_loss, quantized, perplexity, encodings, distances, idx, losses = outs[:7]
        assert len(outs) == 11 and outs[7] is None and outs[8] is None and outs[9] is None
        (vq_loss * float(g['g_loss']) + (quantized * gq).sum()).backward()
        # which rows may legitimately differ: fp64 top-2 gap below 1e-6 relative (reported, not compared)
        flat = vqo.rows_from_bdt(g[f'z{s}'])
        _, near, _ = vqo.assign(flat, W_before)
        ref_idx = g[f'idx{s}'].reshape(-1)
        got_idx = idx.cpu().numpy().reshape(-1)
        assert idx.dtype == torch.int64 and tuple(idx.shape) == (B * T, 1)
        mism = got_idx != ref_idx
        assert not np.any(mism & ~near), 'index mismatch outside near-tie rows'
        n_excluded += int(near.sum())
        if mism.any():      # a near-tie row flipped: every downstream quantity differs legitimately
            pytest.skip('near-tie rows flipped (%d); reported, not compared' % int(mism.sum()))
        ste_scale = max(np.abs(g[f'z{s}']).max(), np.abs(g[f'quantized{s}']).max())
        assert rel_err(quantized.detach().cpu().numpy(), g[f'quantized{s}'], ste_scale) < TOL
        assert rel_err(vq_loss.item(), g[f'vq_loss{s}']) < TOL
        assert abs(losses['vq_loss'] - float(g[f'vq_loss{s}'])) <= TOL * abs(float(g[f'vq_loss{s}']))
        assert rel_err(perplexity.item(), g[f'perplexity{s}']) < TOL
        assert rel_err(outs[10].cpu().numpy(), g[f'concat{s}']) < TOL
        assert rel_err(z.grad.cpu().numpy(), g[f'grad_z{s}']) < TOL
        if s == 0:
            assert np.array_equal(encodings.cpu().numpy(), g['encodings0'])
            assert tuple(encodings.shape) == (B, T, K) and tuple(distances.shape) == (B, T, K)
            d_scale = float(np.max(np.sum(flat.astype(np.float64) ** 2, 1)) + np.max(np.sum(W_before.astype(np.float64) ** 2, 1)))
            assert rel_err(distances.cpu().numpy(), g['distances0'], d_scale) < TOL   # error scales with |x|^2 + |e|^2
            # the returned distances are the ones the indices were taken from
            assert np.array_equal(distances.view(-1, K).argmin(1).cpu().numpy(), got_idx)
        if is_ema:
            assert rel_err(vq._ema_cluster_size.cpu().numpy(), g[f'cs{s + 1}']) < TOL
            assert rel_err(vq._ema_w.detach().cpu().numpy(), g[f'ema_w{s + 1}']) < TOL
            assert rel_err(vq._embedding.weight.detach().cpu().numpy(), g[f'W{s + 1}']) < TOL
            assert vq._embedding.weight.grad is None and vq._ema_w.grad is None
        else:
            gE = vq._embedding.weight.grad.cpu().numpy()
            assert rel_err(gE, g[f'grad_E{s}']) < TOL
            vq._embedding.weight.grad = None
            with torch.no_grad():   # the golden run moved the codebook with plain SGD between steps
                vq._embedding.weight -= 0.1 * _t(g[f'grad_E{s}'], dev)
            assert rel_err(vq._embedding.weight.detach().cpu().numpy(), g[f'W{s + 1}']) < TOL
    vq.eval()
    with torch.no_grad():
        if is_ema:
            vq._embedding.weight.copy_(_t(g[f'W{int(g["steps"])}'], dev))
        outs = vq(_t(g['z_eval'], dev), compute_distances_if_possible=False)
    assert np.array_equal(outs[5].cpu().numpy().reshape(-1), g['eval_idx'].reshape(-1))
    assert rel_err(outs[0].item(), g['eval_vq_loss']) < TOL
    assert rel_err(outs[2].item(), g['eval_perplexity']) < TOL
    assert rel_err(outs[10].cpu().numpy(), g['eval_concat']) < TOL
    sc = max(np.abs(g['z_eval']).max(), np.abs(g['eval_quantized']).max())
    assert rel_err(outs[1].cpu().numpy(), g['eval_quantized'], sc) < TOL


def test_true_ties_resolve_to_lowest_index():
    dev = _dev()
    g = load_golden('vq_ema_k44_d64_b2_t24_dup')
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB
    W = _t(g['W0'], dev)
    z = _t(g['z0'], dev)
    ws = ops.vq_workspace(44, 64, dev)
    idx, stats = ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws)
    idx = idx.cpu().numpy()
    assert np.array_equal(idx, g['idx0'].reshape(-1))
    assert not np.any(idx == 22) and not np.any(idx == 43)   # duplicates of code 1 never win


SWEEP = [  # (K, D, B, T, layout-name)  sizes the oracle finishes in seconds
    (44, 64, 2, 24, 'bdt'), (44, 64, 16, 96, 'bdt'), (44, 64, 256, 96, 'bdt'), (29, 64, 7, 33, 'bdt'),
    (10, 2, 2, 24, 'bdt'), (10, 2, 64, 96, 'bdt'), (100, 64, 5, 24, 'bdt'), (512, 64, 32, 96, 'bdt'),
    (1000, 64, 8, 24, 'bdt'), (4096, 64, 4, 96, 'bdt'), (44, 64, 1, 1000, 'flat'), (44, 48, 3, 50, 'bdt'),
    (7, 5, 3, 11, 'bdt'), (1, 64, 2, 24, 'bdt'), (44, 64, 1, 1, 'bdt'), (64, 128, 4, 24, 'bdt'),
    # B >= 4 D: the element-wise kernels walk (B, D, T) in the blocked order (16 batch items x 8 frames per warp); B not a
    # multiple of 16 and an odd number of 4-frame groups exercise its masks
    (44, 64, 260, 20, 'bdt'), (29, 64, 300, 12, 'bdt'), (10, 2, 64, 8, 'bdt'),
    # D = 64, B % 64 == 0, B >= 256: the tiled element-wise kernels (indices staged in shared memory); partial frame block
    (44, 64, 320, 40, 'bdt'), (29, 64, 256, 8, 'bdt'),
    # flat rows, N >= 4096: tiled forward element-wise kernel, partial last tile
    (44, 64, 1, 5000, 'flat'), (29, 64, 1, 4100, 'flat'),
]


@pytest.mark.parametrize('K,D,B,T,lay', SWEEP)
@pytest.mark.parametrize('ema', [True, False])
def test_vq_ops_match_oracle(K, D, B, T, lay, ema):
    """C-ABI calls vs the numpy oracle on seeded inputs: indices bit-exact outside near-ties, statistics, EMA state,
    quantised output, loss, perplexity and gradients within 1e-5."""
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
    rng = np.random.RandomState(K * 1000 + D * 10 + B)
    W = rng.randn(K, D).astype(np.float32) if ema else rng.uniform(-1 / K, 1 / K, (K, D)).astype(np.float32)
    trained = (K * D + B) % 2 == 0
    if lay == 'flat':
        N = B * T
        rows = (W[rng.randint(0, K, N)] + 0.1 * rng.randn(N, D)).astype(np.float32) if trained \
            else rng.randn(N, D).astype(np.float32)
        z_dev = _t(rows, dev)
        layout = LAYOUT_FLAT_ND
    else:
        N = B * T
        if trained:
            rows = (W[rng.randint(0, K, N)] + 0.1 * rng.randn(N, D)).astype(np.float32)
            z = vqo.bdt_from_rows(rows, B, D, T)
        else:
            z = rng.randn(B, D, T).astype(np.float32)
        z_dev = _t(z, dev)
        layout = LAYOUT_BDT_AS_DTB
        rows = vqo.rows_from_bdt(z)
    beta, decay, eps = 0.25, 0.99, 1e-5
    cs0 = rng.rand(K).astype(np.float32) * 3
    ew0 = rng.randn(K, D).astype(np.float32)
    # ---- oracle (fp64 truth; index search fp32) ----
    idx_o, near, _ = vqo.assign(rows, W)
    counts_o, dw_o = vqo.code_stats(rows.astype(np.float64), idx_o, K)
    # ---- CUDA ----
    Wd = _t(W, dev)
    ws = ops.vq_workspace(K, D, dev)
    dmin2 = torch.empty(N, 2, device=dev)
    dist = torch.empty(N, K, device=dev)
    idx, stats = ops.vq_assign(z_dev, Wd, layout, ws, dmin2=dmin2, distances=dist)
    got = idx.cpu().numpy()
    mism = got != idx_o
    assert not np.any(mism & ~near), '%d index mismatches outside near-ties' % int((mism & ~near).sum())
    # d = (|x|^2 + |e|^2) - 2 x.e cancels when x is close to a code: its fp32 rounding error scales with the operands
    # (|x|^2 + |e|^2), not with d itself -- the reference's own fp32 distances carry the same error
    d_scale = float(np.max(np.sum(rows.astype(np.float64) ** 2, 1)) + np.max(np.sum(W.astype(np.float64) ** 2, 1)))
    assert rel_err(dist.cpu().numpy(), vqo.distances_fp64(rows, W), d_scale) < TOL
    d2 = dmin2.cpu().numpy()
    assert np.all(d2[:, 0] <= d2[:, 1])
    assert np.allclose(d2[:, 0], dist.min(1).values.cpu().numpy(), rtol=0, atol=0)
    if mism.any():
        pytest.skip('near-tie rows flipped (%d of %d); reported, not compared' % (int(mism.sum()), N))
    st = stats.cpu().numpy()
    assert np.array_equal(st[:K], counts_o.astype(np.float32))          # cluster counts bit-exact
    assert rel_err(st[K:].reshape(K, D), dw_o) < TOL
    Wq = W.astype(np.float64)
    if ema:
        csd, ewd = _t(cs0, dev), _t(ew0, dev)
        ops.vq_ema_update(csd, ewd, Wd, stats, decay, eps)
        cs_o, ew_o, Wq = vqo.ema_update(cs0, ew0, counts_o, dw_o, decay, eps, np.float64)
        assert rel_err(csd.cpu().numpy(), cs_o) < TOL
        assert rel_err(ewd.cpu().numpy(), ew_o) < TOL
        assert rel_err(Wd.cpu().numpy(), Wq) < TOL
    q_rows = torch.empty(N, D, device=dev)
    out, sc = ops.vq_quantize(z_dev, idx, Wd, layout, ws, stats[:K], N, beta, q_rows=q_rows)
    q_o = Wq[idx_o]
    diff = q_o - rows
    e_latent = np.mean(diff * diff)
    sc = sc.cpu().numpy()
    assert rel_err(sc[1], e_latent) < TOL
    assert rel_err(sc[3], beta * e_latent) < TOL and rel_err(sc[4], (1 + beta) * e_latent) < TOL
    assert rel_err(sc[2], vqo.perplexity(counts_o, N)) < TOL
    assert rel_err(q_rows.cpu().numpy(), q_o) < TOL
    ste_o = rows + (q_o - rows)
    out_rows = out.cpu().numpy() if lay == 'flat' else vqo.rows_from_bdt(out.cpu().numpy())
    assert rel_err(out_rows, ste_o, max(np.abs(rows).max(), np.abs(q_o).max())) < TOL
    # ---- backward ----
    g_np = rng.randn(*z_dev.shape).astype(np.float32)
    gl = torch.tensor([1.5], device=dev)
    gz = ops.vq_backward(_t(g_np, dev), gl, 2.0 * beta / (N * D), z_dev, idx, Wd, layout).cpu().numpy()
    g_rows = g_np if lay == 'flat' else vqo.rows_from_bdt(g_np)
    gz_rows = gz if lay == 'flat' else vqo.rows_from_bdt(gz)
    gz_o = g_rows + 1.5 * beta * 2.0 * (rows - q_o) / (N * D)
    assert rel_err(gz_rows, gz_o) < TOL
    if not ema:
        gE = ops.vq_grad_codebook(stats, Wd, gl, 2.0 / (N * D)).cpu().numpy()
        gE_o = 1.5 * (2.0 / (N * D)) * (counts_o[:, None] * Wq - dw_o)
        assert rel_err(gE, gE_o) < TOL


def test_vq_full_size_properties():
    """BASELINE sizes (N = 2^20 rows, too big for the oracle): size-independent properties."""
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB
    K, D, B, T = 44, 64, 8192, 128
    N = B * T
    gen = torch.Generator(device=dev).manual_seed(1)
    W = torch.randn(K, D, device=dev, generator=gen)
    z = torch.randn(B, D, T, device=dev, generator=gen)
    ws = ops.vq_workspace(K, D, dev)
    idx, stats = ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    counts = stats[:K]
    assert float(counts.sum()) == N                                             # every row counted exactly once
    assert torch.equal(counts, torch.bincount(idx, minlength=K).float())        # counts bit-exact vs the indices
    rows = z.permute(1, 2, 0).contiguous().view(-1, D)                           # checker only (torch reference layout)
    # linearity: sum_k dw_k == sum of all rows; a zero-mean sum cancels, so the fp32 error is relative to sum |x|
    l1 = float(rows.double().abs().sum(0).max())
    assert rel_err(stats[K:].view(K, D).double().sum(0).cpu().numpy(), rows.double().sum(0).cpu().numpy(), l1) < 1e-6
    out, sc = ops.vq_quantize(z, idx, W, LAYOUT_BDT_AS_DTB, ws, counts, N, 0.25)
    q = W[idx]
    out_rows = out.permute(1, 2, 0).contiguous().view(-1, D)
    assert float((out_rows - q).abs().max()) < 1e-5 * float(q.abs().max()) * 4
    # idempotence: quantised rows map to the same codes, with zero loss
    idx2, _ = ops.vq_assign(q.view(D, T, B).permute(2, 0, 1).contiguous(), W, LAYOUT_BDT_AS_DTB, ws)
    assert torch.equal(idx2, idx)
    # distances really are minimal: compare against a blockwise fp64 check on a sample
    sel = torch.randint(0, N, (4096,), device=dev)
    d = torch.cdist(rows[sel].double(), W.double()) ** 2
    best = d.min(1).values
    chosen = d.gather(1, idx[sel].view(-1, 1)).view(-1)
    assert float(((chosen - best) / best).max()) < 1e-5


def test_cpu_tensor_raises():
    _dev()
    from vq_vae_speech_b200.vector_quantizer import VectorQuantizerEMA
    vq = VectorQuantizerEMA(44, 64, 0.25, 0.99, 'cpu')
    with pytest.raises(RuntimeError):
        vq(torch.randn(2, 64, 24))


TC_CASES = [  # (K, D, B, T, layout, data)
    (44, 64, 2, 24, 'bdt', 'randn'), (44, 64, 64, 96, 'bdt', 'randn'), (44, 64, 256, 96, 'bdt', 'trained'),
    (44, 64, 3, 17, 'bdt', 'randn'), (29, 64, 7, 33, 'bdt', 'trained'), (44, 64, 1, 100000, 'flat', 'randn'),
    (256, 64, 16, 96, 'bdt', 'randn'), (100, 32, 9, 50, 'bdt', 'randn'), (64, 32, 4, 24, 'flat', 'trained'),
    (44, 64, 8, 96, 'bdt', 'near_dup'), (1, 64, 2, 24, 'bdt', 'randn'), (44, 64, 1, 777, 'flat', 'near_dup'),
    # large codebooks: streamed tensor-core distance GEMM (vq_search_large_kernel)
    (512, 64, 32, 96, 'bdt', 'randn'), (1000, 64, 8, 24, 'bdt', 'trained'), (4096, 64, 16, 96, 'bdt', 'randn'),
    (4096, 64, 1, 5000, 'flat', 'near_dup'), (300, 32, 3, 50, 'bdt', 'randn'), (513, 64, 1, 129, 'flat', 'trained'),
    # a few enormous codes (what the first EMA steps do to never-used codes): the per-code error bound must not
    # push every row into the exact path, and indices must still agree
    (44, 64, 8, 96, 'bdt', 'huge'), (4096, 64, 8, 96, 'bdt', 'huge'),
    # flat rows, K <= 48: the streaming engine (TMA tiles as raw tf32 operands, vq_assign_tma_kernel)
    (29, 64, 1, 5000, 'flat', 'trained'), (48, 64, 1, 4097, 'flat', 'randn'), (1, 64, 1, 300, 'flat', 'randn'),
    (44, 64, 1, 100, 'flat', 'near_dup'), (17, 64, 1, 33000, 'flat', 'huge'), (44, 64, 1, 1 << 20, 'flat', 'trained'),
    # (B, D, T) rows with B % 64 == 0, K <= 48: the same streaming engine fed by the cp.async gather.  Tile shapes:
    # 16 frames x 8 batch groups with a partial last frame block; 8 x 16; 32 x 4 with 4 of 32 frames in the last block
    (44, 64, 512, 47, 'bdt', 'randn'), (29, 64, 1024, 24, 'bdt', 'trained'), (44, 64, 256, 100, 'bdt', 'near_dup'),
    (48, 64, 2048, 32, 'bdt', 'huge'), (1, 64, 256, 33, 'bdt', 'randn'),
    # batch groups not divisible by 4: 64 frames x 2 groups, 128 frames x 1 group
    (44, 64, 384, 96, 'bdt', 'trained'), (31, 64, 192, 200, 'bdt', 'randn'),
]


@pytest.mark.parametrize('K,D,B,T,lay,data', TC_CASES)
def test_vq_tensor_core_engine_matches_cuda_core_engine(K, D, B, T, lay, data):
    """The tcgen05 search (3xTF32 scores + exact fp32 re-check of near-ties) must return the SAME indices and counts as
    the exact-fp32 CUDA-core search, including true ties and rows whose two best codes are almost equidistant, and both
    must match the oracle outside fp64 near-ties."""
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
    rng = np.random.RandomState(K + D + B + T)
    W = rng.randn(K, D).astype(np.float32)
    N = B * T
    if data == 'near_dup' and K > 4:
        W[K // 2] = W[1]                              # exact duplicate: lowest index must win
        W[K - 1] = W[2] * (1 + 1e-7)                  # almost-duplicate: gap far below the TF32 score error
        W[3] = W[2] + 1e-6 * rng.randn(D).astype(np.float32)
    if data == 'huge':
        W[K // 3: K // 3 + max(K // 8, 2)] *= 3e4
    if data in ('randn', 'huge'):
        rows = rng.randn(N, D).astype(np.float32)
    else:
        rows = (W[rng.randint(0, K, N)] + 0.1 * rng.randn(N, D)).astype(np.float32)
    if lay == 'flat':
        z, layout = _t(rows, dev), LAYOUT_FLAT_ND
    else:
        z, layout = _t(vqo.bdt_from_rows(rows, B, D, T), dev), LAYOUT_BDT_AS_DTB
    Wd = _t(W, dev)
    ws = ops.vq_workspace(K, D, dev)
    try:
        ops.vq_set_engine('cuda_core')
        idx_cc, st_cc = ops.vq_assign(z, Wd, layout, ws)
        idx_cc, st_cc = idx_cc.clone(), st_cc.clone()
        ops.vq_set_engine('tensor_core')
        idx_tc, st_tc = ops.vq_assign(z, Wd, layout, ws)
    finally:
        ops.vq_set_engine('auto')
    assert torch.equal(idx_tc, idx_cc), '%d rows differ between the engines' % int((idx_tc != idx_cc).sum())
    assert torch.equal(st_tc[:K], st_cc[:K])
    assert rel_err(st_tc[K:].cpu().numpy(), st_cc[K:].cpu().numpy()) < TOL
    # vs the oracle: an fp32 evaluation of (|x|^2 + |e|^2) - 2 x.e resolves gaps down to ~1e-6 of its OPERANDS
    # (|x|^2 + |e|^2), not of the (possibly tiny) distance -- the near-duplicate codes built above sit below that, so the
    # near-tie exclusion is taken relative to the operand scale here
    idx_o, near, gap_rel = vqo.assign(rows, W)
    d64 = vqo.distances_fp64(rows, W)
    part = np.partition(d64, 1, axis=1) if K > 1 else np.concatenate([d64, d64 + 1e9], 1)
    scale = (rows.astype(np.float64) ** 2).sum(1) + (W.astype(np.float64) ** 2).sum(1).max()
    near_fp32 = (part[:, 1] - part[:, 0]) < 1e-6 * scale
    mism = idx_tc.cpu().numpy() != idx_o
    assert not np.any(mism & ~(near | near_fp32))
    assert float(st_tc[:K].sum()) == N


def test_vq_streaming_engine_fuzz_matches_cuda_core_engine():
    """Random (K, N, data) draws through the streaming engine (vq_assign_tma_kernel: flat rows, K <= 48) against the exact
    fp32 CUDA-core search: identical indices and counts, dw within 1e-5, and bit-identical results on a second run."""
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
    rng = np.random.RandomState(2024)
    D = 64
    try:
        for case in range(40):
            K = int(rng.randint(1, 49))
            N = int(rng.choice([1, 31, 127, 128, 129, 1000, 4096, 20011, 70000]))
            kind = rng.choice(['randn', 'trained', 'tiny_codes', 'dup', 'scaled'])
            W = rng.randn(K, D).astype(np.float32)
            if kind == 'tiny_codes':
                W *= 0.05                                  # an EMA-shrunk codebook: |e| << |x|
            if kind == 'dup' and K > 2:
                W[K - 1] = W[0]
                W[K // 2] = W[0] * (1 + 1e-7)
            if kind == 'scaled':
                W *= 300.0
            if kind == 'trained':
                rows = (W[rng.randint(0, K, N)] + 0.05 * rng.randn(N, D)).astype(np.float32)
            else:
                rows = (rng.randn(N, D) * (300.0 if kind == 'scaled' else 1.0)).astype(np.float32)
            z, Wd = _t(rows, dev), _t(W, dev)
            ws = ops.vq_workspace(K, D, dev)
            ops.vq_set_engine('cuda_core')
            i_cc, s_cc = ops.vq_assign(z, Wd, LAYOUT_FLAT_ND, ws)
            i_cc, s_cc = i_cc.clone(), s_cc.clone()
            ops.vq_set_engine('tensor_core')
            i_tc, s_tc = ops.vq_assign(z, Wd, LAYOUT_FLAT_ND, ws)
            i_tc, s_tc = i_tc.clone(), s_tc.clone()
            i_t2, s_t2 = ops.vq_assign(z, Wd, LAYOUT_FLAT_ND, ws)
            tag = 'case %d: K=%d N=%d %s' % (case, K, N, kind)
            assert torch.equal(i_tc, i_cc), tag + ': %d rows differ' % int((i_tc != i_cc).sum())
            assert torch.equal(s_tc[:K], s_cc[:K]), tag
            assert rel_err(s_tc[K:].cpu().numpy(), s_cc[K:].cpu().numpy()) < TOL, tag
            assert torch.equal(i_tc, i_t2) and torch.equal(s_tc, s_t2), tag + ': not deterministic'
    finally:
        ops.vq_set_engine('auto')


def test_elementwise_kernel_variants_agree_bitwise():
    """The tiled / blocked element-wise kernels (staged indices) compute exactly what the grid-stride kernels compute:
    quantised output and grad_z bit-identical, the loss (a differently ordered sum) within 1e-6."""
    import os
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
    gen = torch.Generator(device=dev).manual_seed(3)
    K, D = 44, 64
    W = torch.randn(K, D, device=dev, generator=gen)
    ws = ops.vq_workspace(K, D, dev)
    one = torch.ones(1, device=dev)
    envs = ('VQS_EW_FLAT_TILE', 'VQS_EW_NO_TILE', 'VQS_EW_NO_BLK')
    saved = dict((k, os.environ.get(k)) for k in envs)

    def run(z, layout, env):
        for k in envs:
            os.environ.pop(k, None)
        os.environ.update(env)
        idx, stats = ops.vq_assign(z, W, layout, ws)
        N = idx.numel()
        out, sc = ops.vq_quantize(z, idx, W, layout, ws, stats[:K], N, 0.25)
        g = torch.randn(z.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        gz = ops.vq_backward(g, one, 2 * 0.25 / (N * D), z, idx, W, layout)
        return out.clone(), sc[:5].clone(), gz.clone()       # slots 5-7 of the scalar block are unused

    try:
        z = torch.randn(9001, D, device=dev, generator=gen)
        ref = run(z, LAYOUT_FLAT_ND, {'VQS_EW_FLAT_TILE': '0'})
        for mode in ('1', '2'):
            got = run(z, LAYOUT_FLAT_ND, {'VQS_EW_FLAT_TILE': mode})
            assert torch.equal(got[0], ref[0]) and torch.equal(got[2], ref[2])
            assert rel_err(got[1].cpu().numpy(), ref[1].cpu().numpy()) < 1e-6
        z = torch.randn(320, D, 44, device=dev, generator=gen)
        ref = run(z, LAYOUT_BDT_AS_DTB, {'VQS_EW_NO_TILE': '1', 'VQS_EW_NO_BLK': '1'})
        for env in ({}, {'VQS_EW_NO_TILE': '1'}):
            got = run(z, LAYOUT_BDT_AS_DTB, env)
            assert torch.equal(got[0], ref[0]) and torch.equal(got[2], ref[2])
            assert rel_err(got[1].cpu().numpy(), ref[1].cpu().numpy()) < 1e-6
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def test_vq_streaming_engine_bdt_fuzz_matches_cuda_core_engine():
    """The reference's own row layout through the streaming engine (vq_assign_tma_kernel<BDT>: rows gathered from the
    (B, 64, T) tensor by cp.async, B % 64 == 0): identical indices and counts as the exact fp32 CUDA-core search, dw within
    1e-5, bit-identical on a second run, and the indices sit at the reference's row positions (ema.py:101-106)."""
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB
    rng = np.random.RandomState(77)
    D = 64
    try:
        for case in range(24):
            K = int(rng.randint(1, 49))
            Q = int(rng.choice([4, 8, 12, 16, 20, 32, 64]))
            T = int(rng.choice([8, 9, 16, 24, 31, 32, 33, 47, 64, 96, 191]))
            B = 64 * Q
            kind = rng.choice(['randn', 'trained', 'tiny_codes', 'dup'])
            W = rng.randn(K, D).astype(np.float32)
            if kind == 'tiny_codes':
                W *= 0.05
            if kind == 'dup' and K > 2:
                W[K - 1] = W[0]
                W[K // 2] = W[0] * (1 + 1e-7)
            N = B * T
            if kind == 'trained':
                rows = (W[rng.randint(0, K, N)] + 0.05 * rng.randn(N, D)).astype(np.float32)
            else:
                rows = rng.randn(N, D).astype(np.float32)
            z, Wd = _t(vqo.bdt_from_rows(rows, B, D, T), dev), _t(W, dev)
            ws = ops.vq_workspace(K, D, dev)
            ops.vq_set_engine('cuda_core')
            i_cc, s_cc = ops.vq_assign(z, Wd, LAYOUT_BDT_AS_DTB, ws)
            i_cc, s_cc = i_cc.clone(), s_cc.clone()
            ops.vq_set_engine('tensor_core')
            i_tc, s_tc = ops.vq_assign(z, Wd, LAYOUT_BDT_AS_DTB, ws)
            i_tc, s_tc = i_tc.clone(), s_tc.clone()
            i_t2, s_t2 = ops.vq_assign(z, Wd, LAYOUT_BDT_AS_DTB, ws)
            tag = 'case %d: K=%d B=%d T=%d %s' % (case, K, B, T, kind)
            assert torch.equal(i_tc, i_cc), tag + ': %d rows differ' % int((i_tc != i_cc).sum())
            assert torch.equal(s_tc[:K], s_cc[:K]), tag
            assert rel_err(s_tc[K:].cpu().numpy(), s_cc[K:].cpu().numpy()) < TOL, tag
            assert torch.equal(i_tc, i_t2) and torch.equal(s_tc, s_t2), tag + ': not deterministic'
            if case < 8 and kind in ('randn', 'trained'):   # against the oracle on the reference's rows
                idx_o, near, _ = vqo.assign(rows, W)
                mism = i_tc.cpu().numpy() != idx_o
                assert not np.any(mism & ~near), tag
    finally:
        ops.vq_set_engine('auto')


@pytest.mark.parametrize('case', ['evaltables_noema_k44_d64_b2_t24', 'evaltables_noema_k10_d2_b3_t17'])
def test_eval_mode_distance_tables_match_reference(case):
    """Eval-mode forward with compute_distances_if_possible=True: the three pairwise distance tables of
    vector_quantizer.py:108-127 (one vqs_pairwise_l2 launch each instead of O(N^2) Python torch.dist calls) against the
    reference's own output; also for the EMA class, where the reference raises NameError."""
    dev = _dev()
    from vq_vae_speech_b200.vector_quantizer import VectorQuantizer, VectorQuantizerEMA
    from conftest import load_golden
    g = load_golden(case)
    K, D = int(g['K']), int(g['D'])
    z = _t(g['z'], dev)
    for vq in (VectorQuantizer(K, D, 0.25, dev), VectorQuantizerEMA(K, D, 0.25, 0.99, dev)):
        vq = vq.to(dev).eval()
        with torch.no_grad():
            vq._embedding.weight.copy_(_t(g['W'], dev))
        outs = vq(z, compute_distances_if_possible=True)
        assert np.array_equal(outs[5].cpu().numpy().reshape(-1), g['idx'].reshape(-1))
        for slot, key in ((7, 'encoding_distances'), (8, 'embedding_distances'), (9, 'frames_vs_embedding_distances')):
            assert tuple(outs[slot].shape) == g[key].shape, key
            assert rel_err(outs[slot].cpu().numpy(), g[key]) < TOL, key
        assert rel_err(outs[10].cpu().numpy(), g['concat']) < TOL
        assert all(o is None for o in vq(z, compute_distances_if_possible=False)[7:10])
        assert all(o is None for o in vq.train()(z)[7:10])


def test_pairwise_l2_large_matches_oracle():
    """combinations / product ordering at a size where the triangular index needs the exact integer correction."""
    dev = _dev()
    from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND, LAYOUT_BDT_AS_DTB
    rng = np.random.RandomState(5)
    B, D, T, K = 7, 64, 33, 29
    z = rng.randn(B, D, T).astype(np.float32)
    W = rng.randn(K, D).astype(np.float32)
    enc, emb, fve = vqo.eval_distance_tables(z, W)
    zd, Wd = _t(z, dev), _t(W, dev)
    assert rel_err(ops.pairwise_l2(zd, LAYOUT_BDT_AS_DTB, D).cpu().numpy(), enc.reshape(-1)) < TOL
    assert rel_err(ops.pairwise_l2(Wd, LAYOUT_FLAT_ND, D).cpu().numpy(), emb) < TOL
    assert rel_err(ops.pairwise_l2(zd, LAYOUT_BDT_AS_DTB, D, Wd).cpu().numpy(), fve.reshape(-1)) < TOL


@pytest.mark.parametrize('layout_name,shape', [('flat', (8192, 64)), ('flat', (300, 64)), ('bdt', (256, 64, 32)),
                                               ('bdt', (3, 64, 17)), ('bdt', (5, 2, 24))])
def test_gather_forward_and_backward_with_losses_equal_the_exact_pair(layout_name, shape):
    """The captured training step runs the bottleneck as vq_gather (q = W[idx], z not read) + vq_backward_loss (grad_z and the
    forward's losses in one sweep over z).  Against the exact pair vq_quantize + vq_backward on the same inputs: the gathered
    value is bit-equal to W[idx] and within an ulp of max(|x|, |q|) of the straight-through value fl(x + fl(q - x)); grad_z is
    bit-identical; SSE / e_latent / perplexity / vq_loss agree to 1e-6 (a different summation order of the same terms)."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
    dev = torch.device('cuda:0')
    layout = LAYOUT_FLAT_ND if layout_name == 'flat' else LAYOUT_BDT_AS_DTB
    D = shape[1]
    K = 44 if D == 64 else 10
    g = torch.Generator(device=dev).manual_seed(sum(shape))
    z = torch.randn(*shape, device=dev, generator=g)
    W = torch.randn(K, D, device=dev, generator=g)
    gq = torch.randn(*shape, device=dev, generator=g)
    one = torch.ones(1, device=dev)
    ws = ops.vq_workspace(K, D, dev)
    idx, stats = ops.vq_assign(z, W, layout, ws)
    N = idx.numel()
    beta, coef = 0.25, 2 * 0.25 / (N * D)
    q_exact, sc_exact = ops.vq_quantize(z, idx, W, layout, ws, stats[:K], N, beta)
    gz_exact = ops.vq_backward(gq, one, coef, z, idx, W, layout)
    sc_exact = sc_exact.clone()
    q = ops.vq_gather(idx, W, layout, shape)
    gz, sc = ops.vq_backward_loss(gq, one, coef, z, idx, W, layout, ws, stats[:K], N, beta)
    # the gathered value is W[idx] exactly, in z's layout
    rows = W[idx]                                                       # (N, D) in the reference's row order
    if layout_name == 'flat':
        want = rows
    else:
        B, _, T = shape
        want = rows.reshape(D, T, B).permute(2, 0, 1).contiguous()     # inverse of permute(1, 2, 0).view(-1, D)
    assert torch.equal(q, want)
    tol = 2.4e-7 * torch.maximum(z.abs(), q.abs()) + 1e-30
    assert bool(((q - q_exact).abs() <= tol).all())
    assert torch.equal(gz, gz_exact)
    for i in range(5):
        a, b = float(sc[i]), float(sc_exact[i])
        assert abs(a - b) <= 1e-6 * abs(b) + 1e-12, (i, a, b)
