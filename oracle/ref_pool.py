"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

The reference's own tick loop (``optimax_rogue/server/main.py:110-113``: ``game_state.on_tick()`` then
``updater.update(game_state, p1_move, p2_move)``, ``optimax_rogue/logic/updater.py:76-162``) over a batch of
independent games, spread over every host core with ``multiprocessing`` -- the CPU arm of bench.py
(BASELINE.md section 4). The reference modules are imported unmodified from ``/root/reference`` or from
the copy ``oracle/make_ref.py`` ships to the GPU box; both players are the reference's ``RandomBot``
(``optimax_rogue_bots/randombot.py:20-21``), natively seeded, games restart when they end (what the
batched side calls auto-reset).

A "step" is what it is on the GPU side: ONE tick of EVERY game of the batch. Worker w owns
``games / workers`` games for the whole run and ticks each of them once per step.
"""
import multiprocessing as mp
import os
import random
import sys
import time

import numpy as np

from . import ref_harness as rh

def _worker(conn, wid, games_per_worker, width, height, max_ticks, seed_base):
    """Owns ``games_per_worker`` reference games; every 'step' message ticks each of them once."""
    ref = rh.load_reference()
    sys.stdout = open(os.devnull, 'w')          # belt and braces: load_reference already silences the updater's prints
    random.seed(seed_base + wid)
    np.random.seed((seed_base + wid) % (2 ** 32))
    dgen = ref.worldgen.EmptyDungeonGenerator(width, height)
    start = ref.worldgen.TogetherGameStartGenerator(dgen)
    strat = ref.updater.DungeonDespawningStrategy.Unreachable
    in_progress = ref.updater.UpdateResult.InProgress
    b1, b2 = ref.randombot.RandomBot(1), ref.randombot.RandomBot(2)
    games = [[start.setup_game(), ref.updater.Updater(dgen, strat, max_ticks)] for _ in range(games_per_worker)]
    conn.send(len(games))
    while True:
        if conn.recv() != 'step':
            break
        for slot in games:                       # server/main.py:110-113, once per game
            gs, upd = slot
            gs.on_tick()
            res, _ = upd.update(gs, b1.move(gs), b2.move(gs))
            if res != in_progress:               # the episode is over: start the next one (auto-reset)
                slot[0] = start.setup_game()
                slot[1] = ref.updater.Updater(dgen, strat, max_ticks)
        conn.send(len(games))
    conn.close()


class ReferencePool:
    """``ReferencePool(games).step()`` ticks every game once, on all host cores; returns the ticks done."""

    def __init__(self, games, width=60, height=10, max_ticks=1000, seed_base=1000, workers=None):
        if not rh.reference_available():
            raise RuntimeError('no reference tree (neither /root/reference nor oracle/_ref)')
        self.workers = int(workers or os.cpu_count() or 1)
        self.per_worker = max(1, games // self.workers)
        self.games = self.per_worker * self.workers
        ctx = mp.get_context('fork')
        self.procs, self.conns = [], []
        for w in range(self.workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, w, self.per_worker, width, height, max_ticks, seed_base), daemon=True)
            p.start()
            b.close()
            self.procs.append(p)
            self.conns.append(a)
        for c in self.conns:
            c.recv()                              # every worker has built its games

    def step(self):
        for c in self.conns:
            c.send('step')
        return sum(c.recv() for c in self.conns)

    def close(self):
        for c in self.conns:
            try:
                c.send('stop')
                c.close()
            except OSError:
                pass
        for p in self.procs:
            p.join(timeout=10)


def timed_run(games, steps=None, seconds=None, warmup=1, **kw):
    """Ticks ``games`` reference games in lockstep for ``steps`` steps or about ``seconds`` seconds.
    Returns dict(value = game-ticks/s, cores, games, steps, ticks, elapsed)."""
    pool = ReferencePool(games, **kw)
    try:
        for _ in range(max(warmup, 1)):
            pool.step()
        t0 = time.perf_counter()
        ticks = n = 0
        while True:
            ticks += pool.step()
            n += 1
            el = time.perf_counter() - t0
            if (steps is not None and n >= steps) or (steps is None and el >= seconds):
                break
        return {'value': ticks / el, 'cores': pool.workers, 'games': pool.games, 'steps': n, 'ticks': ticks,
                'elapsed': el, 'python': sys.version.split()[0], 'numpy': np.__version__}
    finally:
        pool.close()
