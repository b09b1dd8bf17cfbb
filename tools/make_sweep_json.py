"""Turns tools/kbench.py log lines (tools/profile_round3.sh: gpurun_out/w_sweep.log) into profiles/r02_batch_sweep.json."""
import json, re, sys
rows, section = [], ''
for line in open(sys.argv[1]):
    line = line.strip()
    if line.startswith('==='):
        section = line.strip('= ').strip()
        continue
    m = re.match(r'overlap=(\d) path_flags=(\d+) tpc=(\d+) isolate=(\d) games=(\d+) batches=(\d+) us/step=([\d.]+)', line)
    if not m:
        continue
    ov, pf, tpc, iso, games, nb, us = int(m[1]), int(m[2]), int(m[3]), int(m[4]), int(m[5]), int(m[6]), float(m[7])
    rows.append({'section': section, 'overlap_ticks': bool(ov), 'tiles_per_cta_override': tpc, 'isolated': bool(iso),
                 'same_state_every_step': nb == 1, 'games_per_launch': games, 'rotating_batches': nb, 'us_per_step': us,
                 'game_ticks_per_s': float(f'{games / us * 1e6:.4g}'), 'alg_GBps_61B': round(61 * games / us / 1e3),
                 'frac_of_measured_hbm_peak': round(61 * games / us / 1e3 / 6548.2, 3)})
json.dump({'what': 'k_step_pipe (orx_step, ruleset R0, 60x10 levels on device, auto-reset, uniform random commands), one launch per step in a CUDA graph of '
                   '400 steps (200 at >= 2^20 games), best of 3 replays, CUDA events, one B200, final build of round 2 (tools/profile_round3.sh). '
                   'overlap_ticks = throughput mode (ORX_PATH_TILE_FLAGS: consecutive launches ordered run by run, a run = the tiles of one CTA); '
                   'false = grid-wait mode. isolated = an ordinary tiny kernel between consecutive ticks (what a policy network in the loop does); '
                   'same_state_every_step = one L2-resident state instead of rotating batches larger than the L2.',
           'peak_GBps': 6548.2, 'rows': rows}, open(sys.argv[2], 'w'), indent=1)
print(len(rows), 'rows')
