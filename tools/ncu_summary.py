#!/usr/bin/env python
"""Summarise an ``ncu --set full --import-source on`` capture (read here, on the CPU box) into the
markdown kept under profiles/:  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.md

Prints the headline metrics (duration, DRAM bytes, pipe utilisation, stall reasons), the SASS
opcode mix with its share of the warp-stall samples, and the sample share of the code regions
delimited by the first and the last DMMA of the kernel (-lineinfo, which build.py passes, also maps
the ncu source page to csrc/*.cuh).
"""

import csv
import io
import subprocess
import sys
from collections import Counter

HEADLINE = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_active.avg',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def page(report, name):
    out = subprocess.run(['ncu', '-i', report, '--page', name, '--csv'], check=True,
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    report = sys.argv[1]
    raw = page(report, 'raw')
    header, units = raw[0], raw[1]
    for launch in raw[2:]:
        values = dict(zip(header, launch))
        unit = dict(zip(header, units))
        print('## {}\n'.format(values.get('Kernel Name', '?')))
        print('| metric | value | unit |\n|---|---|---|')
        for key in HEADLINE:
            if key in values:
                print('| {} | {} | {} |'.format(key, values[key], unit[key]))
        print('\nWarp stall reasons (warps stalled per issue-active cycle):\n')
        stalls = [(float(values[k]), k.split('issue_stalled_')[1].split('_per_')[0])
                  for k in header if k.startswith('smsp__average_warps_issue_stalled_')
                  and k.endswith('_per_issue_active.ratio')]
        print(', '.join('{} {:.2f}'.format(n, v) for v, n in sorted(stalls, reverse=True)[:8]))
        print()
    src = page(report, 'source')
    hdr, data = src[1], src[2:]
    i_src, i_samp, i_exec = hdr.index('Source'), hdr.index('# Samples'), hdr.index(
        'Instructions Executed')
    total_s = sum(int(r[i_samp]) for r in data) or 1
    total_e = sum(int(r[i_exec]) for r in data) or 1
    mix_e, mix_s = Counter(), Counter()
    for r in data:
        words = r[i_src].split()
        op = (words[1] if words[0].startswith('@') else words[0]).split('.')[0]
        mix_e[op] += int(r[i_exec])
        mix_s[op] += int(r[i_samp])
    print('SASS opcode mix (first launch in the report): {} warp instructions, {} stall samples\n'
          .format(total_e, total_s))
    print('| opcode | warp instructions | share | share of stall samples |\n|---|---|---|---|')
    for op, count in mix_e.most_common(12):
        print('| {} | {} | {:.1f}% | {:.1f}% |'.format(op, count, 100.0 * count / total_e,
                                                        100.0 * mix_s[op] / total_s))
    # regions: SASS address ranges between the first and last DMMA = contraction loops
    if not any('DMMA' in r[i_src] for r in data):
        return
    first = next(i for i, r in enumerate(data) if 'DMMA' in r[i_src])
    last = max(i for i, r in enumerate(data) if 'DMMA' in r[i_src])
    bar = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[i_src]]

    def share(lo, hi):
        return (100.0 * sum(int(r[i_samp]) for r in data[lo:hi]) / total_s,
                100.0 * sum(int(r[i_exec]) for r in data[lo:hi]) / total_e)

    print('\n| SASS region | stall samples | warp instructions |\n|---|---|---|')
    print('| before the first DMMA (tile setup + occupation phase) | {:.1f}% | {:.1f}% |'.format(
        *share(0, first)))
    print('| first..last DMMA (contraction k-loops incl. row-dot between them) | {:.1f}% | {:.1f}% |'
          .format(*share(first, last + 1)))
    print('| after the last DMMA (row-dot tail, shuffle reduction, chunk fetch, tile barrier) | '
          '{:.1f}% | {:.1f}% |'.format(*share(last + 1, len(data))))
    barrier = sum(int(r[hdr.index('stall_barrier')]) for r in data)
    print('\nSamples stalled on the block barrier: {:.1f}% ({} BAR.SYNC sites)'.format(
        100.0 * barrier / total_s, len(bar)))


if __name__ == '__main__':
    main()
