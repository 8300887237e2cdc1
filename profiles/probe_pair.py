"""A/B probe of the CTA-pair (cta_group::2) GEMM kernels: run once per setting of VQS_GEMM_PAIR, writes the outputs of the
bench-shape conv / dgrad-like / wgrad GEMMs and their CUDA-event timings; `compare` checks the two dumps bit for bit.
    VQS_GEMM_PAIR=0 python profiles/probe_pair.py run gpurun_out/pair0.pt
    VQS_GEMM_PAIR=1 python profiles/probe_pair.py run gpurun_out/pair1.pt
    python profiles/probe_pair.py compare gpurun_out/pair0.pt gpurun_out/pair1.pt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(path):
    from vq_vae_speech_b200 import functional as F, ops
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    out = {}
    times = {}
    for prec in ('3xtf32', 'tf32'):
        ops.set_precision(prec)
        for (B, C, L, k) in ((64, 768, 47, 3), (64, 768, 48, 3), (64, 768, 24, 3), (64, 768, 48, 1), (8, 256, 33, 3)):
            x = torch.randn(B, C, L, device=dev)
            w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
            b = torch.randn(C, device=dev)
            res = torch.randn(B, C, L, device=dev)
            mask_out = torch.empty(B, C, L, dtype=torch.uint8, device=dev)
            y = torch.empty(B, C, L, device=dev)
            y2 = torch.empty(B, C, L, device=dev)
            dW = torch.empty_like(w)
            ws = F._wgrad_ws(C, C, k, B, L, dev)
            A = F.gemm_weight(w, 'conv_fwd')
            pad = k // 2
            tag = '%s_B%d_C%d_L%d_k%d' % (prec, B, C, L, k)
            times[tag + '_fwd_epi'] = timeit(lambda: F.conv1d_forward(x, A, b, 1, pad, out=y, relu=True, mask_out=mask_out, add_post=res))
            times[tag + '_fwd'] = timeit(lambda: F.conv1d_forward(x, A, b, 1, pad, out=y2))
            times[tag + '_wgrad'] = timeit(lambda: F.conv1d_wgrad(y2, x, dW, 1, pad, ws))
            out[tag] = (y.cpu(), mask_out.cpu(), y2.cpu(), dW.cpu())
            flops = 2.0 * C * C * k * B * L
            print('%-34s fwd+epi %.4f ms  fwd %.4f ms (%.1f TFLOP/s)  wgrad %.4f ms (%.1f TFLOP/s)' % (
                tag, times[tag + '_fwd_epi'], times[tag + '_fwd'], flops / times[tag + '_fwd'] / 1e9,
                times[tag + '_wgrad'], flops / times[tag + '_wgrad'] / 1e9), flush=True)
    torch.save({'out': out, 'times': times}, path)


def compare(p0, p1):
    a, b = torch.load(p0), torch.load(p1)
    bad = 0
    for tag in a['out']:
        for i, (u, v) in enumerate(zip(a['out'][tag], b['out'][tag])):
            same = torch.equal(u, v)
            if not same:
                bad += 1
                d = (u.double() - v.double()).abs().max().item()
                print('MISMATCH', tag, i, 'max abs diff', d, 'scale', u.double().abs().max().item())
    for t in a['times']:
        print('%-40s %.4f -> %.4f ms  (x%.2f)' % (t, a['times'][t], b['times'][t], a['times'][t] / b['times'][t]))
    print('bitwise identical' if bad == 0 else '%d tensors differ' % bad)
    return bad


if __name__ == '__main__':
    if sys.argv[1] == 'run':
        run(sys.argv[2])
    else:
        sys.exit(1 if compare(sys.argv[2], sys.argv[3]) else 0)
