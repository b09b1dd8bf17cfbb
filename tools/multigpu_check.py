"""Two (or more) real GPUs: each rank steps its shard of one global batch with orx_rollout, the stats
vector is summed with NCCL, rank 0 replays the whole batch alone and must get identical planes and
stats (shard invariance on real devices). Run:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from optimax_rogue_b200 import SimConfig
from optimax_rogue_b200.game.state import BatchedGameState
from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator
from optimax_rogue_b200.parallel import gather_stats, shard_range

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
os.environ['NCCL_DEBUG'] = 'WARN'
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
N, T = 1 << 20, 200
cfg = SimConfig(max_ticks=150, seed=0x0A11CE, auto_reset=True)
upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, 150, auto_reset=True)
start, count = shard_range(N, rank, world)
gs = BatchedGameState(cfg, count, dev, game_id_base=start)
reset_games(gs)
stats = upd.rollout(gs, 2, 1, T)
gather_stats(stats)                                   # NCCL all-reduce of the 64-byte counter vector
planes = {name: getattr(gs, name) for name in ('pos', 'hp', 'depth', 'stairs', 'tick', 'episode', 'status')}
if rank == 0:
    full = BatchedGameState(cfg, N, dev, game_id_base=0)
    reset_games(full)
    full_stats = upd.rollout(full, 2, 1, T)
    assert torch.equal(full_stats, stats), (full_stats, stats)
    for name, t in planes.items():
        assert torch.equal(t, getattr(full, name)[start:start + count]), name
for r in range(1, world):                             # the other shards, sent to rank 0 for comparison
    for name in sorted(planes):
        if rank == r:
            dist.send(planes[name].contiguous().view(torch.uint8), dst=0)      # NCCL moves bytes
        elif rank == 0:
            s, c = shard_range(N, r, world)
            want = getattr(full, name)[s:s + c].contiguous().view(torch.uint8)
            buf = torch.empty_like(want)
            dist.recv(buf, src=r)
            assert torch.equal(buf, want), (r, name)
dist.barrier()
if rank == 0:
    print(f'multi-GPU check ok: {world} ranks x {count} games x {T} ticks == 1 rank x {N} games; '
          f'stats {stats.tolist()}')
dist.destroy_process_group()
