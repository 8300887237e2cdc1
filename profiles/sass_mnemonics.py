"""Instruction mnemonics per kernel of the shipped library: the evidence that the tcgen05 / TMA / cp.async / multimem / PDL paths
are what the binary contains.  Runs without a GPU.

    python profiles/sass_mnemonics.py > profiles/<name>_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'vq-vae-speech_b200', 'csrc', 'libvqs_b200.so')
COLS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'ACQBULK', 'LDGMC', 'FMNMX3', 'FMNMX', 'REDG', 'FFMA',
        'LDS', 'STS']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return out


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            counts[cur][m.group(1)] += 1
    names = demangle(order)
    print('# cuobjdump -sass vq-vae-speech_b200/csrc/libvqs_b200.so (sm_100a): instruction mnemonics per kernel.  UTCHMMA = tcgen05.mma,')
    print('# UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, LDGSTS = cp.async,')
    print('# SYNCS = mbarrier, ACQBULK = griddepcontrol.wait (PDL), LDGMC = multimem.ld_reduce (NVLS; multimem.st is an STG...SYS to the multicast address), REDG = red.global, FMNMX3 = 3-input min/max')
    print('%-84s' % 'kernel' + ''.join('%9s' % c for c in COLS))
    for mangled, name in zip(order, names):
        short = name.replace('void ', '').replace('(anonymous namespace)::', '')
        short = re.sub(r'_GLOBAL__N__\w+::', '', short)
        short = re.sub(r'\(.*$', '', short).replace('vqs::', '')
        c = counts[mangled]
        agg = dict((col, c.get(col, 0)) for col in COLS)
        print('%-84s' % short[:84] + ''.join('%9d' % agg[col] for col in COLS))


if __name__ == '__main__':
    main()
