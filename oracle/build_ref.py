"""Recipe for oracle/_ref: the reference's own hot-path files, UNMODIFIED, copied from /root/reference/src -- TEST /
BASELINE INFRASTRUCTURE ONLY (only tests/, __graft_entry__ and bench.py's CPU legs may use oracle/).

The reference is pure Python (no build step); "building" it means placing the files of the path where the GPU box can
import them: /root/reference does not exist there, oracle/_ref/ travels with the snapshot (git-ignored, NOT
gpurun-ignored -- the same treatment the base contract gives a pip-installed baseline/_ref).  Nothing under oracle/_ref/
is ever committed or edited; bench.py times exactly these files (`cpu_baseline.kind` = "reference").

    python oracle/build_ref.py            # no-op with a message when /root/reference is absent (GPU box)

Files (the import closure of models.convolutional_vq_vae + experiments.convolutional_trainer minus plotting):
  models/{convolutional_vq_vae,convolutional_encoder,deconvolutional_decoder,vector_quantizer,vector_quantizer_ema}.py
  modules/{conv1d_builder,conv_transpose1d_builder,residual,residual_stack,jitter}.py
  speech_utils/global_conditioning.py, error_handling/{console_logger,color_print}.py
  experiments/{base_trainer,convolutional_trainer}.py    (ConvolutionalTrainer.iterate is the timed call)
`evaluation/gradient_stats.py` is NOT copied: it imports matplotlib (absent here and on the GPU box) and is only reached
when record_codebook_stats is on; oracle/ref_harness.py registers an empty stand-in module for that one import.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('VQS_REFERENCE_SRC', '/root/reference/src')
DST = os.path.join(HERE, '_ref', 'src')

FILES = [
    'models/__init__.py', 'models/convolutional_vq_vae.py', 'models/convolutional_encoder.py',
    'models/deconvolutional_decoder.py', 'models/vector_quantizer.py', 'models/vector_quantizer_ema.py',
    'modules/__init__.py', 'modules/conv1d_builder.py', 'modules/conv_transpose1d_builder.py', 'modules/residual.py',
    'modules/residual_stack.py', 'modules/jitter.py',
    'speech_utils/__init__.py', 'speech_utils/global_conditioning.py',
    'error_handling/__init__.py', 'error_handling/console_logger.py', 'error_handling/color_print.py',
    'experiments/__init__.py', 'experiments/base_trainer.py', 'experiments/convolutional_trainer.py',
]


def build(verbose=False):
    """Copies FILES from the reference tree into oracle/_ref/src and writes MANIFEST (sha256 per file).  Returns the
    destination, or None when the reference tree is not present (then an existing oracle/_ref is used as is)."""
    if not os.path.isdir(SRC):
        if verbose:
            print('reference tree %s not present: keeping %s' % (SRC, DST if os.path.isdir(DST) else '(nothing)'))
        return DST if os.path.isdir(DST) else None
    lines = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, 'rb') as f:
            lines.append('%s  %s' % (hashlib.sha256(f.read()).hexdigest(), rel))
    with open(os.path.join(HERE, '_ref', 'MANIFEST'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    if verbose:
        print('copied %d reference files to %s' % (len(FILES), DST))
    return DST


if __name__ == '__main__':
    build(verbose=True)
    sys.exit(0)
