"""Round 2: turns the artefacts tools/profile_round3.sh (final build; profile_round2.sh earlier in the round) brought back in gpurun_out/ into committed summaries under profiles/.
    python tools/summarize_profiles2.py"""
import json, os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import summarize_profiles as sp

ROOT, OUT = sp.ROOT, os.path.join(sp.ROOT, 'gpurun_out')
tag = 'r02'
print(sp.summarize_launch_list(os.path.join(OUT, 'launches_r02.csv'), tag))
shutil.copyfile(os.path.join(OUT, 'launches_r02.csv'), os.path.join(sp.PROF, 'r02_launch_list.csv'))
traffic = {}
for name, games, what in (('pipe2p20t', 1 << 20, 'throughput mode (ORX_PATH_TILE_FLAGS; the mode of bench.py\'s value leg), 2^20 games per launch'),
                          ('pipe2p20', 1 << 20, 'grid-wait mode (default), 2^20 games per launch'),
                          ('pipe2p17', 1 << 17, 'throughput mode (ORX_PATH_TILE_FLAGS), 2^17 games per launch'),
                          ('pipe2p24', 1 << 24, 'grid-wait mode, 2^24 games per launch: 512 MB of planes, 4x the L2, so the written planes reach DRAM inside the profiled window')):
    rep = sp.summarize_report(os.path.join(OUT, f'prof_{name}_r02.ncu-rep'), tag, f'k_step_{name}')
    rd = [sp.to_bytes(*l['dram__bytes_read.sum']) for l in rep]
    wr = [sp.to_bytes(*l['dram__bytes_write.sum']) for l in rep]
    us = [float(l['gpu__time_duration.sum'][0]) / (1e3 if l['gpu__time_duration.sum'][1] == 'ns' else 1.0) for l in rep]
    traffic[name] = {'what': what, 'kernel': rep[0]['kernel'], 'launches_profiled': len(rep), 'games_per_launch': games,
                     'dram_bytes_read_per_launch': sum(rd) / len(rd), 'dram_bytes_write_per_launch': sum(wr) / len(wr),
                     'dram_bytes_per_game': (sum(rd) + sum(wr)) / len(rd) / games, 'profiled_duration_us': us,
                     'registers_per_thread': rep[0].get('launch__registers_per_thread', [None])[0],
                     'grid_size': rep[0].get('launch__grid_size', [None])[0]}
main = traffic['pipe2p20']
out = {'kernel': main['kernel'], 'launches_profiled': main['launches_profiled'],
       'dram_bytes_read_per_launch': main['dram_bytes_read_per_launch'], 'dram_bytes_write_per_launch': main['dram_bytes_write_per_launch'],
       'dram_bytes_per_launch': main['dram_bytes_read_per_launch'] + main['dram_bytes_write_per_launch'],
       'note': 'round 2 final build, ncu --set full --clock-control none, 2^20 games per launch (grid-wait mode). Under ncu each launch runs alone with a '
               'cold L2, so the written planes are still dirty in the 126 MB L2 when the launch ends: the read side (31 B/game = 29 B planes + 2 B '
               'commands) is the algorithmic read traffic; the write side is verified by large_batch_check (2^24 games, planes 4x the L2)',
       'source': 'prof_pipe2p20_r02.ncu-rep (summary: profiles/r02_k_step_pipe2p20_ncu_full.json)',
       'large_batch_check': dict(traffic['pipe2p24'], note='read + write DRAM bytes per game against the 61 B algorithmic figure; source prof_pipe2p24_r02.ncu-rep'),
       'throughput_mode_2p17': traffic['pipe2p17'], 'throughput_mode_2p20': traffic['pipe2p20t']}
json.dump(out, open(os.path.join(sp.PROF, 'roofline_traffic.json'), 'w'), indent=1)
for f in ('kernel_trace_2p20_r02.csv', 'kernel_trace_2p20_r02.txt', 'kernel_trace_2p17_r02.csv', 'kernel_trace_2p17_r02.txt'):
    shutil.copyfile(os.path.join(OUT, f), os.path.join(sp.PROF, f.replace('kernel_trace', 'r02_kernel_trace').replace('_r02.', '.')))
print(json.dumps({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != 'kernel'}) for k, v in out.items() if k != 'kernel'}, indent=1)[:3000])
