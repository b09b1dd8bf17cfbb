"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_profiles.py <round-tag> <launches.csv> <prof_pipe.ncu-rep> [<other.ncu-rep> ...]
  (<prof_pipe.ncu-rep> as captured by tools/profile_round.sh)
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, 'profiles')

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
STALLS = 'smsp__average_warps_issue_stalled_'


def raw_rows(rep):
    """Raw page of a report: from <rep>.raw.csv when the box exported it (gpurun brings back at most 64 MiB, the
    reports themselves are 10-20 MB each), else from the .ncu-rep."""
    if os.path.exists(rep + '.raw.csv'):
        out = open(rep + '.raw.csv').read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(val, unit):
    m = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(val) * m.get(unit, 1)


def summarize_report(rep, tag, name, first=0, count=None):
    hdr, units, rows = raw_rows(rep)
    rows = rows[first:first + count] if count else rows[first:]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows:
        d = {'kernel': r[idx['Kernel Name']]}
        for k in KEYS:
            if k in idx:
                d[k] = (r[idx[k]], units[idx[k]])
        d['stalls_per_issue'] = {h[len(STALLS):].replace('_per_issue_active.ratio', ''): float(r[i])
                                 for h, i in idx.items() if h.startswith(STALLS) and r[i] not in ('', 'no data')}
        launches.append(d)
    path = os.path.join(PROF, f'{tag}_{name}_ncu_full.json')
    with open(path, 'w') as f:
        json.dump(launches, f, indent=1)
    return launches


def summarize_launch_list(path_csv, tag):
    rows = list(csv.DictReader(l for l in open(path_csv) if not l.startswith('==')))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r['Kernel Name']
        agg[k][0] += 1
        agg[k][1] += float(r['Metric Value'])
    total = sum(v[1] for v in agg.values())
    out = os.path.join(PROF, f'{tag}_launch_list_summary.csv')
    with open(out, 'w') as f:
        f.write('kernel,launches,total_us,avg_us,share_of_profiled_time\n')
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'"{k}",{v[0]},{v[1] / 1e3:.1f},{v[1] / 1e3 / v[0]:.2f},{v[1] / total:.4f}\n')
    return out


def main():
    tag, launches_csv, step_rep = sys.argv[1:4]
    os.makedirs(PROF, exist_ok=True)
    print(summarize_launch_list(launches_csv, tag))
    # tools/profile_targets.py order: 2 ticks, 2 observes, 2 fused tick+observe (launches 4..9 of the run)
    step = summarize_report(step_rep, tag, 'k_step_pipe', 0, 2)
    summarize_report(step_rep, tag, 'observe_pipe', 2, 2)
    summarize_report(step_rep, tag, 'step_observe_pipe', 4, 2)
    rd = [to_bytes(*l['dram__bytes_read.sum']) for l in step]
    wr = [to_bytes(*l['dram__bytes_write.sum']) for l in step]
    traffic = {'kernel': step[0]['kernel'], 'launches_profiled': len(step),
               'dram_bytes_read_per_launch': sum(rd) / len(rd), 'dram_bytes_write_per_launch': sum(wr) / len(wr),
               'dram_bytes_per_launch': (sum(rd) + sum(wr)) / len(rd),
               'note': 'ncu --set full, 2^20 games per launch; under ncu each launch runs alone with a cold L2, '
                       'so the written planes are still dirty in the 126 MB L2 when the launch ends: the read '
                       'side (31 B/game = 29 B planes + 2 B commands) equals the algorithmic read bytes, the '
                       '30 B/game of writes reach DRAM after the profiled window',
               'source': os.path.basename(step_rep)}
    with open(os.path.join(PROF, 'roofline_traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1)
    print(traffic)
    for extra in sys.argv[4:]:
        name = os.path.basename(extra).replace('.ncu-rep', '').replace('prof_', '').replace('_' + tag, '')
        rep = summarize_report(extra, tag, name)
        if 'step16m' in name:
            rd16 = [to_bytes(*l['dram__bytes_read.sum']) for l in rep]
            wr16 = [to_bytes(*l['dram__bytes_write.sum']) for l in rep]
            traffic['large_batch_check'] = {
                'games_per_launch': 1 << 24, 'dram_bytes_read_per_launch': sum(rd16) / len(rd16),
                'dram_bytes_write_per_launch': sum(wr16) / len(wr16),
                'bytes_per_game': (sum(rd16) + sum(wr16)) / len(rd16) / (1 << 24),
                'note': '2^24 games per launch (512 MB of planes, 4x the L2): written planes must spill to DRAM '
                        'inside the profiled window, so read+write per game can be compared with the 61 B algorithmic figure'}
            with open(os.path.join(PROF, 'roofline_traffic.json'), 'w') as f:
                json.dump(traffic, f, indent=1)


if __name__ == '__main__':
    main()
