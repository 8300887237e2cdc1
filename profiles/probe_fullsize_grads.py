"""Per-parameter gradient deviations at the benchmarked size (C = 768, B = 64, T = 47): ours (3xtf32 tcgen05 / exact-fp32
CUDA cores) vs the torch-CPU fp32 reference, and both against the same step in fp64 (oracle/torch_port.py in double)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_fullsize_gpu import _Reference
from conftest import rel_err
from vq_vae_speech_b200.convolutional_vq_vae import ConvolutionalVQVAE
from vq_vae_speech_b200.trainer import FusedTrainStep, reference_config
dev = torch.device('cuda:0')
B, T, seed = 64, 47, 1234
cfg = reference_config(decay=0.99, batch_size=B)
torch.manual_seed(seed); np.random.seed(seed)
model0 = ConvolutionalVQVAE(cfg, 'cpu')
sd = {k: v.detach().clone() for k, v in model0.state_dict().items()}
ref, ref64 = _Reference(cfg, sd, seed), _Reference(cfg, sd, seed, exact=True)
x = torch.randn(B, T, 39, generator=torch.Generator().manual_seed(seed))
ref.step(x); ref64.step(x)
rg, rg64 = ref.grads(), ref64.grads()
res = {}
for prec in ('3xtf32', 'fp32'):
    torch.manual_seed(seed); np.random.seed(seed)
    m = ConvolutionalVQVAE(cfg, 'cpu'); m.load_state_dict(sd); m = m.to(dev).train()
    eng = FusedTrainStep(m, B, T, cfg['learning_rate'], use_graph=False, precision=prec)
    eng.step(x); eng.losses()
    res[prec] = {n: g.cpu().numpy() for n, g in eng.gradients().items()}
def l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def outl(a, b, thr=1e-5):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b) / np.abs(b).max()
    return int((d > thr).sum()), d.size, float(np.quantile(d, 0.999))
print('%-52s | max-norm: %9s %9s %9s | L2: %9s %9s %9s | entries > 1e-5 (tc, cc) of n | p99.9 tc cc' % ('parameter', 'ref32-f64', 'tc-f64', 'cc-f64', 'ref32-f64', 'tc-f64', 'cc-f64'))
for n in rg:
    ot, oc = outl(res['3xtf32'][n], rg64[n]), outl(res['fp32'][n], rg64[n])
    print('GRADS %-46s | %9.2e %9.2e %9.2e | %9.2e %9.2e %9.2e | %6d %6d of %8d | %8.1e %8.1e' % (
        n, rel_err(rg[n], rg64[n]), rel_err(res['3xtf32'][n], rg64[n]), rel_err(res['fp32'][n], rg64[n]),
        l2(rg[n], rg64[n]), l2(res['3xtf32'][n], rg64[n]), l2(res['fp32'][n], rg64[n]), ot[0], oc[0], ot[1], ot[2], oc[2]), flush=True)
