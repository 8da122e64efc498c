#!/usr/bin/env python
"""Static SASS statistics of one kernel of a built .so (no GPU needed).

    python tools/sass_stats.py rayoptics_b200/csrc/libb200rt.so 'k_trace_grid_leanILi0ELb1ELb0ELb0'

Prints the instruction count, opcode-class histogram and, for every backward
branch (loop), the address range and its histogram.  A proxy used to compare
kernel variants before spending GPU time: it says nothing about stalls.
"""
import collections
import re
import subprocess
import sys

CLASSES = [('fp64', ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')),
           ('mufu', ('MUFU',)),
           ('mov', ('MOV', 'IMAD.MOV', 'UMOV', 'SEL', 'FSEL', 'PRMT', 'SHFL')),
           ('int', ('IADD', 'IMAD', 'LOP3', 'SHF', 'LEA', 'ISETP', 'IABS', 'I2F', 'F2I', 'UIADD', 'ULOP', 'UISETP',
                    'ULEA', 'USHF', 'UIMAD', 'VIADD', 'R2UR', 'S2R', 'S2UR', 'CS2R', 'PLOP3', 'UPLOP3', 'P2R', 'R2P')),
           ('fp32', ('FFMA', 'FSETP', 'FMUL', 'FADD', 'FMNMX', 'FCHK')),
           ('mem', ('LDS', 'STS', 'LDG', 'STG', 'LDC', 'LDCU', 'LDL', 'STL', 'ATOM', 'RED', 'ULDC')),
           ('branch', ('BRA', 'BSSY', 'BSYNC', 'CALL', 'RET', 'EXIT', 'WARPSYNC', 'BAR', 'BREAK', 'NOP', 'YIELD',
                       'BMOV', 'DEPBAR', 'ERRBAR', 'MEMBAR'))]


def classify(op):
    for name, pre in CLASSES:
        for p in pre:
            if op == p or op.startswith(p + '.') or (p == 'IMAD.MOV' and op.startswith('IMAD.MOV')):
                if name == 'int' and op.startswith('IMAD.MOV'):
                    return 'mov'
                return name
    return 'other'


def functions(so):
    txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    cur, out = None, {}
    for ln in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', ln)
        if m and cur is not None:
            body = m.group(2).strip()
            pred = None
            pm = re.match(r'(@!?U?P\d+)\s+(.*)', body)
            if pm:
                pred, body = pm.group(1), pm.group(2)
            op = body.split()[0]
            out[cur].append((int(m.group(1), 16), op, body, pred))
    return out


def hist(ins):
    h = collections.Counter(classify(op) for _, op, _, _ in ins)
    return ' '.join(f'{k}={h[k]}' for k in ('fp64', 'mufu', 'mov', 'int', 'fp32', 'mem', 'branch', 'other')
                    if h[k]) + f' | total={len(ins)}'


def main():
    so, pat = sys.argv[1], sys.argv[2]
    for name, ins in functions(so).items():
        if pat not in name:
            continue
        print(name)
        print('  all:', hist(ins))
        for a, op, body, pred in ins:
            if op.startswith('BRA'):
                m = re.search(r'0x([0-9a-f]+)', body)
                if m and int(m.group(1), 16) < a:
                    t = int(m.group(1), 16)
                    rng = [i for i in ins if t <= i[0] <= a]
                    print(f'  loop {t:#06x}..{a:#06x} ({pred or "always"}):', hist(rng))
        if len(sys.argv) > 3:
            ops = collections.Counter(op for _, op, _, _ in ins)
            for op, c in ops.most_common(40):
                print(f'    {op:24s} {c}')


if __name__ == '__main__':
    main()
