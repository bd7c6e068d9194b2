#!/usr/bin/env python
"""Opcode census of the shipped kernels: cuobjdump -sass libpwv_b200.so, counted per kernel.

    python tools/sass_census.py > profiles/r2_sass_census.txt

The mnemonics that prove Blackwell-native code (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'parallel-wavenet-vocoder_b200', 'libpwv_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'USETMAXREG', 'SYNCS', 'HMMA', 'FFMA', 'FFMA2', 'MUFU', 'LDGSTS',
        'STL', 'LDL']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r'\(.*', '', name).replace('void pwv::', ''), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            op = m.group(1)
            cur['_total'] += 1
            for k in KEYS:
                if op == k or (k.endswith('MMA') and op.startswith(k)):
                    cur[k] += 1
    print('# SASS opcode census of', os.path.relpath(LIB, ROOT), '(cuobjdump -sass, sm_100a)')
    print('%-64s %7s ' % ('kernel', 'instr') + ' '.join('%8s' % k for k in KEYS))
    tot = collections.Counter()
    for name, c in per.items():
        print('%-64s %7d ' % (name[:64], c['_total']) + ' '.join('%8d' % c[k] for k in KEYS))
        tot.update(c)
    print('%-64s %7d ' % ('TOTAL', tot['_total']) + ' '.join('%8d' % tot[k] for k in KEYS))


if __name__ == '__main__':
    main()
