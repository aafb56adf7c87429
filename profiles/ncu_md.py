"""Summarise one ncu --set full report (any kernel) as a markdown table under profiles/.

    python profiles/ncu_md.py gpurun_out/r2_dwt_reg_model.ncu-rep profiles/r2_dwt_reg_model.md "command line"
"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
    'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
    'sm__cycles_active.min', 'sm__cycles_active.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_read.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
]


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    name = r[hdr.index('Kernel Name')]
    lines = [f'# ncu --set full: `{name}`', '', f'Report: `{rep}`.  Command: `{cmd}`', '',
             '| metric | value | unit |', '|---|---|---|']
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f'| `{k}` | {r[i]} | {units[i]} |')
    stalls = {}
    for i, h in enumerate(hdr):
        if 'warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.05:
                stalls[h.split('issue_stalled_')[1].split('_per_issue')[0]] = v
    lines += ['', 'Warp stall reasons (per issue-active cycle, > 0.05):', '', '| stall | ratio |', '|---|---|']
    lines += [f'| {k} | {v:.3f} |' for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])]
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
