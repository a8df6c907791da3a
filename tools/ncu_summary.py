"""Summarise an ncu report exported with --page raw / --page source (csv): headline metrics, stall mix, opcode mix."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
h, v = rows[0], rows[2]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    for hh, uu, vv in zip(h, rows[1], v):
        if hh == k:
            print("%-80s %s %s" % (k, vv, uu))
print("-- stalls per issued instruction")
for hh, vv in zip(h, v):
    if 'issue_stalled' in hh and 'per_issue_active' in hh and float(vv) > 0.05:
        print("   %-30s %s" % (hh.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), vv))
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
isrc, ix, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[ix]) for r in data)
tots = sum(int(r[isamp]) for r in data)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    op = r[isrc].strip().split()
    if op and op[0].startswith('@'):
        op = op[1:]
    o = op[0].split('.')[0] if op else '?'
    ops[o] += int(r[ix])
    samp[o] += int(r[isamp])
print("-- opcode mix (share of executed warp instructions, share of stall samples); total %d" % tot)
for o, c in ops.most_common(18):
    print("   %-10s %5.1f%%  %5.1f%%" % (o, 100 * c / tot, 100 * samp[o] / max(tots, 1)))
