"""Profiling aid: builds a VARIANT of libvqs_b200.so (extra -D flags on chosen sources) next to the shipped one.

    python profiles/build_variant.py <tag> "<extra nvcc flags>" [source.cu ...]     -> csrc/libvqs_b200.<tag>.so

The other objects are taken from the shipped build; select the variant with VQS_LIB_PATH=<path> (see _lib.load()).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'vq-vae-speech_b200'))
import build as b  # noqa: E402

tag, flags, srcs = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
b.build()
objs = []
procs = []
for src in b.sources():
    base = os.path.basename(src)
    if base in srcs:
        obj = src[:-3] + '.' + tag + '.o'
        procs.append(subprocess.Popen([b.NVCC] + b.ARCH + b.FLAGS + flags + ['-Xcompiler', '-fPIC', '-c', src, '-o', obj]))
    else:
        obj = src[:-3] + '.o'
    objs.append(obj)
for p in procs:
    if p.wait() != 0:
        sys.exit(1)
out = os.path.join(b.CSRC, 'libvqs_b200.%s.so' % tag)
subprocess.check_call([b.NVCC] + b.ARCH + ['-shared', '-o', out] + objs + ['-lcuda'])
print(out)
