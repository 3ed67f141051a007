#!/usr/bin/env python
"""Opcode summary of every kernel in libbags_b200.so (cuobjdump -sass), written to profiles/sass_summary.txt.

    python tools/sass_summary.py [path/to/libbags_b200.so] [out.txt]

The SASS mnemonics that prove a Blackwell-native kernel (B200 profiling guide): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor copies, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, RED / ATOM
with .SYS scope or multimem (LDGMC / STGMC / REDGMC) = peer / multicast traffic of the gradient exchange.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'UTCCP',
       'SYNCS', 'MUFU', 'HMMA', 'RED', 'ATOM', 'LDGMC', 'STGMC', 'REDGMC', 'MEMBAR', 'ACQBULK', 'BAR', 'LDS', 'STS',
       'LDG', 'STG', 'STAS', 'CCTL']


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'balancedgroupsoftmax_b200', 'libbags_b200.so')
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'profiles', 'sass_summary.txt')
    txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
    demangle = {}
    funcs = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)', line)
        if m:
            op = m.group(1)
            funcs[cur]['_total'] += 1
            funcs[cur][op.split('.')[0]] += 1
            if op.startswith(('UTCHMMA', 'UTCQMMA')) and '2CTA' in op:
                funcs[cur]['UTC*MMA.2CTA'] += 1
            if '.SYS' in op and op.split('.')[0] in ('ST', 'LD', 'STG', 'LDG', 'ATOM', 'ATOMG', 'RED', 'MEMBAR'):
                funcs[cur][op.split('.')[0] + '.SYS'] += 1
    names = list(funcs)
    try:
        dm = subprocess.run(['cu++filt'] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    lines = ['# SASS opcode summary of %s (cuobjdump -sass, sm_100a); %d kernels' % (os.path.basename(so), len(funcs)), '']
    tot = collections.Counter()
    for name, c in funcs.items():
        full = demangle.get(name, name)
        cut = full.rfind('>(')
        short = full[:cut + 1] if cut >= 0 else full.split('(')[0]
        short = short.replace('(int)', '').replace('(bool)', '')
        keys = [k for k in KEY + ['UTC*MMA.2CTA', 'ST.SYS', 'LD.SYS', 'ATOM.SYS', 'ATOMG.SYS', 'MEMBAR.SYS'] if c.get(k)]
        lines.append('%-112s %6d instr  %s' % (short.replace('void bags::', '').replace('bags::', '')[:112], c['_total'], '  '.join('%s=%d' % (k, c[k]) for k in keys)))
        tot.update({k: v for k, v in c.items()})
    lines += ['', 'library totals: ' + '  '.join('%s=%d' % (k, tot[k]) for k in KEY + ['UTC*MMA.2CTA'] if tot.get(k))]
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[-3:]))
    print('wrote', out)


if __name__ == '__main__':
    main()
