"""One small launch of every warp-specialised mbarrier / TMEM / TMA kernel family, for compute-sanitizer (one tool per gpurun
call):   compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/sanitize_driver.py
Each result is checked against torch (fp64) so that a sanitizer-clean run is also a correct one."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, functional as F, LAYOUT_FLAT_ND, LAYOUT_BDT_AS_DTB
dev = torch.device('cuda:0')
torch.manual_seed(0)
def rel(a, r):
    return ((a.double() - r).abs().max() / r.abs().max()).item()
ops.set_precision('3xtf32')
# gemm_tc_kernel conv mode: BN = 128, BN = 64 + split-K (few tiles), stride-2 dgrad; wgrad gather kernel; wgrad_tma_kernel
for (B, Cin, Cout, L, k, stride, pad) in ((8, 128, 256, 48, 3, 1, 1), (2, 64, 64, 24, 3, 1, 1), (3, 64, 96, 47, 4, 2, 2), (4, 128, 128, 48, 1, 1, 0)):
    x = torch.randn(B, Cin, L, device=dev); w = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
    y64 = torch.nn.functional.conv1d(x.double(), w.double(), None, stride, pad)
    gy = torch.randn(*y64.shape, device=dev)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    (torch.nn.functional.conv1d(xr, wr, None, stride, pad) * gy.double()).sum().backward()
    ws = F._wgrad_ws(Cout, Cin, k, B, y64.shape[2], dev)
    y = F.conv1d_forward(x, F.gemm_weight(w, 'conv_fwd'), None, stride, pad, splitk_ws=ws)
    dx = F.conv1d_dgrad(gy, F.gemm_weight(w, 'conv_dgrad'), L, stride, pad, splitk_ws=ws)
    dW = torch.empty_like(w)
    F.conv1d_wgrad(gy, x, dW, stride, pad, ws)
    torch.cuda.synchronize()
    assert rel(y, y64) < 1e-5 and rel(dx, xr.grad) < 1e-5 and rel(dW, wr.grad) < 1e-5, (B, Cin, Cout, L, k)
    print('conv ok', B, Cin, Cout, L, k, flush=True)
# VQ: streaming search (TMA flat rows, cp.async (B, 64, T) rows), resident-codebook tcgen05 search, large-codebook search
for (K, shape, layout, eng) in ((44, (8192, 64), LAYOUT_FLAT_ND, 'auto'), (44, (256, 64, 32), LAYOUT_BDT_AS_DTB, 'auto'),
                                (100, (1024, 64), LAYOUT_FLAT_ND, 'tensor_core'), (512, (1024, 64), LAYOUT_FLAT_ND, 'auto')):
    ops.vq_set_engine(eng)
    z = torch.randn(*shape, device=dev); W = torch.randn(K, 64, device=dev)
    ws = ops.vq_workspace(K, 64, dev)
    idx, st = ops.vq_assign(z, W, layout, ws)
    rows = z if layout == LAYOUT_FLAT_ND else z.permute(1, 2, 0).contiguous().view(-1, 64)
    d = (rows.double() ** 2).sum(1, keepdim=True) + (W.double() ** 2).sum(1) - 2 * rows.double() @ W.double().t()
    ref = d.argmin(1)
    bad = (idx != ref)
    if bad.any():      # only fp64 near-ties may differ
        two = d.topk(2, dim=1, largest=False).values
        assert bool((((two[:, 1] - two[:, 0]) / two[:, 0].abs())[bad] < 1e-6).all())
    q = ops.vq_gather(idx, W, layout, shape)
    gz, sc = ops.vq_backward_loss(torch.randn(*shape, device=dev), torch.ones(1, device=dev), 1e-4, z, idx, W, layout, ws, st[:K], idx.numel(), 0.25)
    torch.cuda.synchronize()
    assert torch.equal(st[:K].cpu(), torch.bincount(idx, minlength=K).float().cpu())
    print('vq ok', K, shape, eng, flush=True)
ops.vq_set_engine('auto')
print('sanitize driver done')
