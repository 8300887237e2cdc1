"""BASELINE.json configs[4]: batch-size / utterance-length sweep of the full training step with jitter enabled.

    python profiles/step_sweep.py [out.jsonl]

Runs bench.py once per (per-GPU batch, frames) point (batch 2 -> 256 at T = 47 as in
configurations/experiments_vq44-mfcc39-batch_sizes.json extended to 256; lengths 47 / 95 / 191 frames = 7680 / 15360 /
30720 samples) and keeps, per point, ms/step, utterances/s, frames/s, e2e and the GEMM rates."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POINTS = [(2, 47), (16, 47), (32, 47), (64, 47), (128, 47), (256, 47), (64, 95), (64, 191), (16, 191), (256, 191)]


def main(out):
    with open(out, 'w') as f:
        for batch, frames in POINTS:
            cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--batch', str(batch), '--frames', str(frames), '--jitter',
                   '--skip-cpu', '--skip-vq', '--steps', '20', '--warmup', '5']
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            if r.returncode != 0:
                rec = {'batch': batch, 'frames': frames, 'error': r.stderr[-400:]}
            else:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                kb = j['kernel_breakdown']
                rec = {'batch': batch, 'frames': frames, 'use_jitter': True, 'ms_per_step': round(j['ms_per_step'], 4),
                       'utterances_per_s': round(j['value'], 1), 'frames_per_s': round(j['value'] * frames, 1),
                       'e2e_utterances_per_s': round(j['e2e']['value'], 1),
                       'conv_gemm_tflops': kb['vqs_conv_gemm']['tflops'], 'wgrad_gemm_tflops': kb['vqs_wgrad_gemm']['tflops'],
                       'conv_gemm_ms': kb['vqs_conv_gemm']['ms_per_step'], 'wgrad_gemm_ms': kb['vqs_wgrad_gemm']['ms_per_step'],
                       'flops_per_step': j['flops_per_step'], 'clocks': j['clocks']}
            f.write(json.dumps(rec) + '\n')
            f.flush()
            print(json.dumps(rec), file=sys.stderr)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/step_sweep.jsonl')
