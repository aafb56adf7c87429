"""Turn the scratch ncu outputs under gpurun_out/ into the committed summaries.

    python profiles/summarize_ncu.py gpurun_out/prof_chisq_r1.ncu-rep gpurun_out/launches_r1.csv r1

writes profiles/<tag>_model_chisq.{md,json} and profiles/<tag>_launches.md.
Runs in the authoring container (ncu can read reports without a GPU).
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_static',
    'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__warps_eligible.avg.per_cycle_active',
    'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'sm__cycles_active.min',
    'sm__cycles_active.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sectors_op_read.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
]


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(v)*mult.get(unit, 1)


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    hdr, units, rows = raw(rep)
    r = rows[0]
    name = r[hdr.index('Kernel Name')]
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals[k] = (r[i], units[i])
    stalls = {}
    for i, h in enumerate(hdr):
        if 'warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.05:
                stalls[h.split('issue_stalled_')[1].split('_per_issue')[0]] = v
    dr = to_bytes(*vals['dram__bytes_read.sum']) + to_bytes(*vals['dram__bytes_write.sum'])
    js = {'kernel': name, 'source': f'ncu --set full, {os.path.basename(rep)} ({tag})',
          'dram_bytes_per_launch': dr,
          'duration_us_under_ncu': float(vals['gpu__time_duration.sum'][0]) *
          (1e-3 if vals['gpu__time_duration.sum'][1] == 'ns' else 1.0 if vals['gpu__time_duration.sum'][1] == 'us' else 1e3),
          'fp64_pipe_pct_of_active': float(vals['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'][0]),
          'registers': int(float(vals['launch__registers_per_thread'][0])),
          'grid': int(float(vals['launch__grid_size'][0]))}
    json.dump(js, open(os.path.join(HERE, f'{tag}_model_chisq.json'), 'w'), indent=1)
    with open(os.path.join(HERE, f'{tag}_model_chisq.md'), 'w') as f:
        f.write(f'# {tag}: ncu --set full of the dominant kernel\n\n`{name}`\n\n'
                'Command: `ncu --set full --clock-control none --import-source on -k regex:k_model_chisq '
                '-s 8 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu` (config 2: 4096 chains x 1e5 points, fp64).\n\n'
                '| metric | value | unit |\n|---|---|---|\n')
        for k in KEYS:
            if k in vals:
                f.write(f'| `{k}` | {vals[k][0]} | {vals[k][1]} |\n')
        f.write('\nWarp stall reasons (per issue-active cycle, > 0.05):\n\n| stall | ratio |\n|---|---|\n')
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
            f.write(f'| {k} | {v:.3f} |\n')
    # launch list
    rows = [x for x in csv.reader(open(launches)) if len(x) > 10]
    h = rows[0]
    ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for x in rows[1:]:
        try:
            v = float(x[iv].replace(',', ''))
        except ValueError:
            continue
        if x[iu] == 'us':
            v *= 1e3
        a = agg.setdefault(x[ik][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(HERE, f'{tag}_launches.md'), 'w') as f:
        f.write(f'# {tag}: launch list of `bench.py --steps 5 --warmup 3 --no-cpu`\n\n'
                'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv`. '
                'Times are cold-cache and serialised: read the SHARES.\n\n'
                '| launches | total ns | share | kernel |\n|---:|---:|---:|---|\n')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| {n} | {t:.0f} | {100*t/tot:.2f}% | `{k}` |\n')
    print('wrote', tag)


if __name__ == '__main__':
    main()
