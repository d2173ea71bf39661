#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key raw metrics + SASS stall samples by opcode and by region.
usage: tools/ncu_summary.py report.ncu-rep [block]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for i, h in enumerate(hdr):
        if h in WANT or 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            vals = [r[i] for r in rows[2:]]
            if 'issue_stalled' in h and all(float(v or 0) < 0.15 for v in vals):
                continue
            print(f"{h},{units[i]},{','.join(vals)}")
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    sec = []
    for r in rows[2:]:
        if len(r) < 10:
            break
        sec.append(r)
    tot = sum(int(r[idx['# Samples']]) for r in sec) or 1
    byop, ex = collections.Counter(), collections.Counter()
    st = collections.Counter()
    for r in sec:
        t = r[idx['Source']].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        byop[op] += int(r[idx['# Samples']])
        ex[op] += int(r[idx['Instructions Executed']])
        for k in hdr:
            if k.startswith('stall_') and 'Not Issued' not in k:
                st[k] += int(r[idx[k]] or 0)
    print("# samples by opcode (share, executed):", [(o, round(100 * v / tot, 1), ex[o]) for o, v in byop.most_common(14)])
    print("# stall reasons:", [(k, round(100 * v / tot, 1)) for k, v in st.most_common(8)])
    if B:
        for a in range(0, len(sec), B):
            blk = sec[a:a + B]
            s = sum(int(r[idx['# Samples']]) for r in blk)
            if s > 0.01 * tot:
                ops = collections.Counter((r[idx['Source']].split()[0]).split('.')[0] for r in blk)
                print(f"# sass[{a}:{a + B}] {100 * s / tot:5.1f}% ", ops.most_common(4))


if __name__ == "__main__":
    main()
