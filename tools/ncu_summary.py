#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of metrics we track.
Usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

WANT = [
    ('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs'), ('launch__occupancy_limit_registers', 'occ_lim_regs'),
    ('launch__occupancy_limit_shared_mem', 'occ_lim_smem'), ('launch__occupancy_limit_warps', 'occ_lim_warps'),
    ('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('smsp__inst_executed.sum', 'warp_insts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wavefronts'),
    ('l1tex__data_pipe_lsu_wavefronts.sum', 'lsu_wavefronts'),
    ('smsp__inst_executed_op_local_ld.sum', 'local_ld'), ('smsp__inst_executed_op_local_st.sum', 'local_st'),
    ('smsp__average_warp_latency_issue_stalled_barrier.ratio', 'stall_barrier'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_sb'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_sb'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall_mio'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall_lg'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall_math'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall_wait'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall_not_sel'),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ['# ncu summary of %s' % rep, '']
    for r in rows[2:]:
        ent = []
        for h, short in WANT:
            if h in col:
                v = r[col[h]]
                u = units[col[h]]
                if h == 'Kernel Name':
                    v = v[:60]
                ent.append('%s=%s%s' % (short, v, (' ' + u) if u and short in ('time', 'dram_rd', 'dram_wr') else ''))
        lines.append('- ' + ', '.join(ent))
    out = '\n'.join(lines) + '\n'
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(out)
    print(out)


if __name__ == '__main__':
    main()
