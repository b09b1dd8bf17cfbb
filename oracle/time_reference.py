"""TEST INFRASTRUCTURE -- times the LIVE Python reference (BASELINE.md section 4 / SURVEY.md 8(d) config 1)
on the host cores of the machine it runs on. Only usable where /root/reference exists (the build
container); the GPU box has no reference tree, so bench.py's CPU legs use the C port there and this
script's output is committed as context under profiles/.

    python -m oracle.time_reference [seconds]
"""
import json
import multiprocessing as mp
import os
import platform
import random
import sys
import time

import numpy as np

from . import ref_harness as rh


def worker(args):
    wid, seconds = args
    ref = rh.load_reference()
    dgen = ref.worldgen.EmptyDungeonGenerator(60, 10)
    ticks, episodes, seed = 0, 0, wid * 100000
    t_end = time.perf_counter() + seconds
    devnull = open(os.devnull, 'w')
    sys.stdout = devnull
    while time.perf_counter() < t_end:
        random.seed(seed); np.random.seed(seed % (2**32)); seed += 1
        gs = ref.worldgen.TogetherGameStartGenerator(dgen).setup_game()
        upd = ref.updater.Updater(dgen, ref.updater.DungeonDespawningStrategy.Unreachable, 20000)
        b1, b2 = ref.randombot.RandomBot(1), ref.randombot.RandomBot(2)
        res = ref.updater.UpdateResult.InProgress
        while res == ref.updater.UpdateResult.InProgress and time.perf_counter() < t_end:
            gs.on_tick()
            res, _ = upd.update(gs, b1.move(gs), b2.move(gs))
            ticks += 1
        episodes += 1
    sys.stdout = sys.__stdout__
    return ticks, episodes


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    cores = os.cpu_count()
    t0 = time.perf_counter()
    with mp.Pool(cores) as pool:
        res = pool.map(worker, [(w, seconds) for w in range(cores)])
    el = time.perf_counter() - t0
    ticks = sum(r[0] for r in res)
    t1 = time.perf_counter()
    one = worker((999, min(seconds, 10.0)))
    el1 = time.perf_counter() - t1
    out = {'what': 'live Python reference: Updater.update + GameState.on_tick + 2 x RandomBot.move, natively seeded, '
                   'TogetherGameStartGenerator(EmptyDungeonGenerator(60,10)), Unreachable, max_ticks=20000, stdout discarded',
           'where': 'build container (NOT the GPU box)', 'cores': cores, 'python': platform.python_version(),
           'numpy': np.__version__, 'wall_s': el, 'game_ticks': ticks, 'game_ticks_per_s_all_cores': ticks / el,
           'game_ticks_per_s_one_core': one[0] / el1}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
