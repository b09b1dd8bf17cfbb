"""N>1 host logic on CPU: two gloo ranks shard a batch by global game id, step their shards (the
oracle stands in for the kernels -- tests may do that), reduce the stats vector, and must
reproduce the single-process run bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cport
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.parallel import gather_stats, max_over_ranks, shard_range

N_TOTAL, TICKS = 1001, 120


def _cfg():
    return SimConfig(max_ticks=60, seed=77, auto_reset=True)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    start, count = shard_range(N_TOTAL, rank, world)
    orc = cport.Oracle(_cfg(), count, game_id_base=start)
    orc.reset()
    stats = orc.rollout(1, 2, TICKS)
    t = torch.from_numpy(stats.astype(np.int64))
    gather_stats(t)
    slow = max_over_ranks(float(rank + 1))
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), start=start, count=count, stats=t.numpy(), slow=slow,
             pos=orc.state.pos, hp=orc.state.hp, depth=orc.state.depth, tick=orc.state.tick,
             episode=orc.state.episode)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_gloo_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = cport.Oracle(_cfg(), N_TOTAL, game_id_base=0)
    full.reset()
    full_stats = full.rollout(1, 2, TICKS)
    parts = [np.load(os.path.join(str(tmp_path), f'r{r}.npz')) for r in range(world)]
    for p in parts:
        s, c = int(p['start']), int(p['count'])
        for name in ('pos', 'hp', 'depth', 'tick', 'episode'):
            assert np.array_equal(p[name], getattr(full.state, name)[s:s + c]), name
        assert np.array_equal(p['stats'].astype(np.uint64), full_stats)     # reduced over both ranks
        assert float(p['slow']) == 2.0
