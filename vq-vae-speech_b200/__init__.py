"""vq-vae-speech_b200: B200-native (sm_100a) implementation of the VQ-VAE-Speech training hot path.

Layout
  csrc/                  hand-written CUDA kernels + the C ABI (include/vqs_b200.h) -> libvqs_b200.so
  _lib.py, ops.py        ctypes binding / tensor-level wrappers (torch = device memory + streams only)
  functional.py          conv-like GEMM mappings and autograd Functions
  vector_quantizer.py    VectorQuantizer, VectorQuantizerEMA            (reference src/models/vector_quantizer{,_ema}.py)
  modules.py             Conv1DBuilder, ConvTranspose1DBuilder, Residual, ResidualStack, Jitter  (reference src/modules/)
  convolutional_vq_vae.py ConvolutionalEncoder, DeconvolutionalDecoder, ConvolutionalVQVAE       (reference src/models/)
  trainer.py             fused, CUDA-graph-captured, data-parallel training step (reference convolutional_trainer.py:44-74)

The directory name carries a hyphen (task-mandated); `import vq_vae_speech_b200` (the sibling shim package) loads it.
"""
from ._lib import LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND  # noqa: F401

__all__ = ['LAYOUT_BDT_AS_DTB', 'LAYOUT_FLAT_ND']
