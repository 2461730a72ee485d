"""Turn ncu exports brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.txt
  python tools/summarize_ncu.py full gpurun_out/prof_tc.ncu-rep profiles/r1_pw_tc_ncu.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio')


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [r for r in rows if r[0] == 'ID'][0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    data = [r for r in rows if r[0].isdigit()]
    for r in data:
        name = re.sub(r'\(.*', '', r[ik]).replace('<unnamed>::', '').replace('void ', '')
        a = agg.setdefault(name, [0.0, 0])
        a[0] += float(r[iv].replace(',', ''))
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# {src}: {len(data)} launches, {tot / 1e6:.2f} ms of kernel time (gpu__time_duration.sum, ncu-serialised, cold cache: compare SHARES)\n')
        f.write(f'{"kernel":70s} {"total us":>10s} {"launches":>8s} {"share":>7s}\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f'{k[:70]:70s} {a[0] / 1e3:10.1f} {a[1]:8d} {100 * a[0] / tot:6.1f}%\n')


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
            f.write(f'## {name}  grid {vals[hdr.index("Grid Size")]} block {vals[hdr.index("Block Size")]}\n')
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    f.write(f'{h:90s} {u:16s} {v}\n')
            f.write('\n')




def traffic(dst, *reports):
    """average dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel in the given reports -> JSON for bench.py"""
    import json
    agg = {}
    for src in reports:
        out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        for v in rows[2:]:
            name = re.sub(r'[<(].*', '', v[ik].replace('void ', '').replace('<unnamed>::', ''))
            b = float(v[ir].replace(',', '')) * scale[units[ir]] + float(v[iw].replace(',', '')) * scale[units[iw]]
            a = agg.setdefault(name, [0.0, 0])
            a[0] += b
            a[1] += 1
    with open(dst, 'w') as f:
        json.dump({k: a[0] / a[1] for k, a in agg.items()}, f, indent=1)


if __name__ == '__main__' and sys.argv[1] == 'traffic':
    traffic(sys.argv[2], *sys.argv[3:])
elif __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
