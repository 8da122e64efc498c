#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) into a small markdown file for profiles/.

usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_<kernel>.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
    'dram__bytes_write.sum.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
]


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True)
    return list(csv.reader(out.stdout.splitlines()))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    raw = page(rep, 'raw')
    hdr, units = raw[0], raw[1]
    lines = [f'# ncu summary of `{rep.split("/")[-1]}`', '',
             'captured with `ncu --set full --clock-control none --import-source on` under gpurun '
             '(one B200); numbers under the profiler are for shares and counters, never bench values.', '']
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        lines += [f'## {d.get("Kernel Name", "?")}', '', '| metric | value | unit |', '|---|---|---|']
        for k in KEYS:
            if k in d:
                lines.append(f'| {k} | {d[k]} | {units[hdr.index(k)]} |')
        lines += ['', '| stall reason (warps per issue) | value |', '|---|---|']
        for k in hdr:
            if 'issue_stalled' in k and 'per_issue_active' in k and float(d[k] or 0) > 0.05:
                lines.append(f'| {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")} | {d[k]} |')
        lines.append('')
    src = page(rep, 'source')
    if len(src) > 2:
        h = src[1]
        ai, ei = h.index('Source'), h.index('Instructions Executed')
        c, tot = collections.Counter(), 0
        for r in src[2:]:
            try:
                n = int(r[ei])
            except (ValueError, IndexError):
                continue
            t = re.sub(r'^@!?U?P\d+\s+', '', r[ai].strip())
            op = t.split()[0]
            base = 'IMAD.MOV' if op.startswith('IMAD.MOV') else op.split('.')[0]
            c[base] += n
            tot += n
        lines += ['## dynamic SASS opcode mix (warp instructions, first kernel)', '',
                  '| opcode | executed | share |', '|---|---|---|']
        for k, v in c.most_common(20):
            lines.append(f'| {k} | {v} | {100*v/tot:.1f} % |')
        lines.append(f'| total | {tot} | |')
    open(dst, 'w').write('\n'.join(lines) + '\n')
    print('wrote', dst)
    # per-launch DRAM traffic of the first kernel, for bench.py's roofline.traffic
    if len(sys.argv) > 3:
        import json
        d = dict(zip(hdr, raw[2]))
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tr = sum(float(d[k])*scale[units[hdr.index(k)]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        def pct(k):
            return float(d[k]) if k in d else None
        json.dump({'kernel': d.get('Kernel Name', '?'), 'dram_bytes_per_launch': tr,
                   'fp64_pipe_pct_of_peak': pct('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
                   'issue_active_pct': pct('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                   'dram_throughput_pct': pct('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                   'kernel_us_under_ncu': pct('gpu__time_duration.sum'),
                   'source': rep.split('/')[-1] + ' (ncu --set full, one launch)'},
                  open(sys.argv[3], 'w'), indent=1)
        print('wrote', sys.argv[3])


if __name__ == '__main__':
    main()
