"""K = 4096 search, N = 2^20: time of the shipped kernel, of the cluster-multicast variants and of the phase probes of a
-DVQS_DEBUG build (VQS_TC_DEBUG bits: 8 no scan, 16 no MMAs, 32 no chunk loads, 64 no settlement of near-ties).  One subprocess per variant.

    python profiles/probe_k4096_phases.py            (variants built by profiles/build_variant.py)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'vq-vae-speech_b200', 'csrc')
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
K, D, N = 4096, 64, 1 << 20
g = torch.Generator(device=dev).manual_seed(K)
W = torch.randn(K, D, device=dev, generator=g); z = torch.randn(N, D, device=dev, generator=g)
ws = ops.vq_workspace(K, D, dev)
ops.vq_set_engine('tensor_core')
idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('%%-28s %%.4f ms = %%.1f TFLOP/s algorithmic' %% (os.environ.get('VQS_TAG'), ms, 2.0 * N * D * K / ms / 1e9), flush=True)
''' % ROOT

runs = [('shipped', None, None)]
for tag in sys.argv[1:] or ['cl2', 'cl4']:
    runs.append((tag, 'libvqs_b200.%s.so' % tag, None))
if os.path.exists(os.path.join(CSRC, 'libvqs_b200.dbg.so')):
    for bits in [int(b) for b in os.environ.get("VQS_PROBE_BITS", "0,64,8,16,32,24,40,48,56").split(",")]:
        runs.append(('dbg bits=%d' % bits, 'libvqs_b200.dbg.so', bits))
for tag, lib, bits in runs:
    env = dict(os.environ, VQS_TAG=tag)
    if lib:
        if not os.path.exists(os.path.join(CSRC, lib)):
            continue
        env['VQS_LIB_PATH'] = os.path.join(CSRC, lib)
    if bits is not None:
        env['VQS_TC_DEBUG'] = str(bits)
    subprocess.run([sys.executable, '-c', CHILD], env=env)
