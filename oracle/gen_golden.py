"""TEST INFRASTRUCTURE (oracle) -- generates tests/golden/ from the LIVE reference.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden [--quick]
Everything written here is a function of the unmodified reference code executed under the
injected Philox draw schedule (oracle/ref_harness.py); nothing comes from the oracle or the
CUDA path. Committed outputs:

  tests/golden/digests_stair_vs_random.npy   uint64[10000]  StaircaseBot vs RandomBot, max_ticks 512
  tests/golden/digests_random_vs_random.npy  uint64[10000]  RandomBot vs RandomBot, max_ticks 1000
  tests/golden/digests_meta.json             seeds / config / per-suite result histogram
  tests/golden/traces.npz + cases.json       full per-tick traces of assorted configurations
  tests/golden/truth_table.json              two-player resolution table
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

from . import philox as px
from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
SEED = 0x0A11CE

SUITES = {
    'stair_vs_random': dict(bots=('staircase', 'random'), max_ticks=512, despawn='unreachable'),
    'random_vs_random': dict(bots=('random', 'random'), max_ticks=1000, despawn='unreachable'),
    # both players descend all the time: double descents, DungeonCreated under Unused, spawn retries
    'stair_vs_stair_unused': dict(bots=('staircase', 'staircase'), max_ticks=256, despawn='unused', episodes=4000),
    # SeparatedGameStartGenerator (worldgen.py:91-135): player 1 climbs down towards player 2's level
    'separated_stair_vs_random': dict(bots=('staircase', 'random'), max_ticks=256, despawn='unreachable',
                                      start='separated', p_depths=(0, 6), episodes=3000),
}


def _digest_job(args):
    name, gid = args
    kw = {k: v for k, v in SUITES[name].items() if k != 'episodes'}
    trace, _ = rh.play_episode(SEED, gid, **kw)
    return gid, rh.digest(trace), len(trace) - 1, trace[-1]['result'], trace[-1]['ent'][0][2], trace[-1]['ent'][1][2]


def fixed_map(w=60, h=10, p=0.10, stairs=False, seed=0):
    """Same construction as tests/test_gpu_parity.py:fixed_map (config 2 of BASELINE.json)."""
    rng = np.random.default_rng(seed)
    t = np.full((w, h), 1, np.uint8)
    t[[0, -1], :] = 2
    t[:, [0, -1]] = 2
    inner = rng.random((w - 2, h - 2)) < p
    t[1:-1, 1:-1][inner] = 2
    if stairs:
        g = np.argwhere(t == 1)
        for k in (len(g) // 3, 2 * len(g) // 3):
            t[g[k][0], g[k][1]] = 3
    return t


def trace_cases():
    cases = []
    for gid in range(4):
        cases.append(dict(game_id=gid, bots=('staircase', 'random'), max_ticks=256))
        cases.append(dict(game_id=gid, bots=('random', 'random'), max_ticks=400))
    cases.append(dict(game_id=7, bots=('staircase', 'staircase'), max_ticks=200, despawn='unused'))
    cases.append(dict(game_id=8, bots=('random', 'staircase'), max_ticks=200, despawn='unused'))
    cases.append(dict(game_id=9, bots=('staircase', 'random'), max_ticks=300, start='separated', p_depths=(3, 5)))
    cases.append(dict(game_id=10, bots=('staircase', 'staircase'), max_ticks=300, start='separated', p_depths=(0, 4), despawn='unused'))
    cases.append(dict(game_id=11, bots=('random', 'random'), max_ticks=300, width=4, height=4, hp=3))
    cases.append(dict(game_id=12, bots=('random', 'random'), max_ticks=300, width=5, height=7, hp=(4, 6), damage=(3, 2), armor=(1, 0)))
    cases.append(dict(game_id=13, bots=('random', 'random'), max_ticks=300, width=6, height=4, hp=2))
    cases.append(dict(game_id=14, bots=('random', 'random'), max_ticks=300, fixed='walls'))
    cases.append(dict(game_id=15, bots=('staircase', 'random'), max_ticks=300, fixed='walls_stairs'))
    cases.append(dict(game_id=16, bots=('random', 'staircase'), max_ticks=200, width=8, height=6, hp=50,
                      npcs=[(0, 3, 3, 2), (0, 5, 2, 1), (1, 2, 2, 3)]))
    cases.append(dict(game_id=(1 << 40) + 5, bots=('staircase', 'random'), max_ticks=128, episode=3))
    return cases


def run_case(case):
    kw = dict(case)
    gid = kw.pop('game_id')
    episode = kw.pop('episode', 0)
    fixed = kw.pop('fixed', None)
    if fixed is not None:
        kw['fixed_tiles'] = fixed_map(stairs=(fixed == 'walls_stairs')).astype('int32')
    trace, moves = rh.play_episode(SEED, gid, episode, **kw)
    max_ev = 4 + len(case.get('npcs', ()))
    recs = np.array([rh.record_values(r, max_ev) for r in trace], dtype=np.int64)
    return recs, np.array(moves, dtype=np.uint8)


def truth_table():
    """8 relative placements x 25 move pairs x 2 initiative orders on an open 9x9 room without
    stairs (fixed map), one tick each, from the live reference."""
    tiles = np.full((9, 9), 1, 'int32')
    tiles[[0, -1], :] = 2
    tiles[:, [0, -1]] = 2
    # one game id per initiative order at tick 1
    gids = {}
    g = 0
    while len(gids) < 2:
        j = px.bounded(px.block(SEED, g, 0, px.DOM_TICK, 0, 1)[2], 2)
        gids.setdefault(j, g)
        g += 1
    rows = []
    placements = [(1, 0), (-1, 0), (0, 1), (0, -1), (2, 0), (0, 2), (1, 1), (-1, 1)]
    for (dx, dy) in placements:
        for m1 in range(1, 6):
            for m2 in range(1, 6):
                for j, gid in sorted(gids.items()):
                    p1 = (4, 4)
                    p2 = (4 + dx, 4 + dy)
                    trace, _ = rh.play_episode(SEED, gid, bots=('script', 'script'), scripts=[[m1], [m2]],
                                               fixed_tiles=tiles, place=(p1, p2), limit_ticks=1, max_ticks=None)
                    r = trace[1]
                    rows.append({'placement': [dx, dy], 'm1': m1, 'm2': m2, 'gid': gid,
                                 'first': 'p2' if j == 0 else 'p1',
                                 'ent': [list(e) for e in r['ent']], 'events': [list(e) for e in r['events']],
                                 'result': r['result']})
    return rows


def main():
    quick = '--quick' in sys.argv
    only_new = '--only-new' in sys.argv        # keep committed suites, add the missing ones
    os.makedirs(OUT, exist_ok=True)
    n_ep = 200 if quick else 10000
    meta = {'seed': SEED, 'episodes': n_ep, 'suites': {}}
    meta_path = os.path.join(OUT, 'digests_meta.json')
    if only_new and os.path.exists(meta_path):
        meta = json.load(open(meta_path))
    with mp.Pool(os.cpu_count()) as pool:
        for name, kw in SUITES.items():
            if only_new and name in meta['suites'] and os.path.exists(os.path.join(OUT, f'digests_{name}.npy')):
                continue
            n_suite = min(n_ep, kw.get('episodes', n_ep))
            res = pool.map(_digest_job, [(name, g) for g in range(n_suite)], chunksize=16)
            res.sort()
            dig = np.array([r[1] for r in res], dtype=np.uint64)
            np.save(os.path.join(OUT, f'digests_{name}.npy'), dig)
            ticks = [r[2] for r in res]
            hist = {str(k): int(sum(1 for r in res if r[3] == k)) for k in (2, 3, 4)}
            meta['suites'][name] = dict({k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
                                        bots=list(kw['bots']), episodes=n_suite, total_ticks=int(sum(ticks)),
                                        result_hist=hist, max_depth=int(max(max(r[4], r[5]) for r in res)))
            print(name, meta['suites'][name], flush=True)
    with open(meta_path, 'w') as f:
        json.dump(meta, f, indent=1)
    if only_new:
        return

    cases = trace_cases()
    arrays = {}
    for k, case in enumerate(cases):
        recs, moves = run_case(case)
        arrays[f'case{k}_records'] = recs.astype(np.int32)
        arrays[f'case{k}_moves'] = moves
        print('case', k, case, recs.shape, flush=True)
    np.savez_compressed(os.path.join(OUT, 'traces.npz'), **arrays)
    with open(os.path.join(OUT, 'cases.json'), 'w') as f:
        json.dump(cases, f, indent=1)

    tt = truth_table()
    with open(os.path.join(OUT, 'truth_table.json'), 'w') as f:
        json.dump(tt, f)
    print('truth table rows', len(tt))


if __name__ == '__main__':
    main()
