"""Builds libvqs_b200.so (sm_100a only) in-tree with nvcc.  No torch involved: the library is a plain C-ABI shared object.

    python vq-vae-speech_b200/build.py [--force] [--verbose]
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, 'libvqs_b200.so')
STAMP = os.path.join(CSRC, '.libvqs_b200.stamp')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo'] + os.environ.get('VQS_EXTRA_NVCC_FLAGS', '').split()   # e.g. -DVQS_DEBUG, -DVQS_LARGE_TBUF=2 (profiles/build_variant.py builds such variants next to the shipped library)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    for p in sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + sorted(glob.glob(os.path.join(ROOT, 'include', '*.h'))):
        h.update(os.path.relpath(p, ROOT).encode())     # relative: the repo lives at a different path on the GPU box
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libvqs_b200.so.  Returns the library path."""
    dig = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig

    if not force and fresh():
        return LIB
    # one builder at a time (several ranks import the package concurrently under torch.distributed.run)
    import fcntl
    lock = open(os.path.join(CSRC, '.build.lock'), 'w')
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and fresh():
            return LIB
        return _build_locked(dig, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(dig, verbose):
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):      # GPU box without a toolkit: use the prebuilt library that travelled with the snapshot
            return LIB
        raise RuntimeError('nvcc not found at %s and no prebuilt %s' % (NVCC, LIB))
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + '.o'
        cmd = [NVCC] + ARCH + FLAGS + ['-Xcompiler', '-fPIC', '-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
        if verbose or out.strip():
            sys.stderr.write(out)
    tmp = LIB + '.tmp.%d' % os.getpid()
    cmd = [NVCC] + ARCH + ['-shared', '-o', tmp] + objs + ['-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    os.replace(tmp, LIB)          # atomic: a concurrent loader never sees a half-written library
    with open(STAMP, 'w') as f:
        f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
