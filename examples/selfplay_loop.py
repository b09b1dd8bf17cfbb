#!/usr/bin/env python
"""A self-play loop on the batched updater: observation -> (toy) policy -> commands -> tick.

Everything stays on the GPU: `observe` writes int16[N,2,12] per-player views, a small torch network
maps them to a distribution over the five Move codes for each player, `update` advances all N
games with auto-reset, and the result codes give the terminal rewards. This is the loop the
reference runs over TCP with one game per server process (optimax_rogue/server/main.py:110-113,
optimax_rogue_bots/main.py:122-124), here for N games per launch.

    python examples/selfplay_loop.py --games 65536 --ticks 200
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from optimax_rogue_b200 import _abi
from optimax_rogue_b200.logic.updater import BatchedUpdater, DungeonDespawningStrategy, UpdateResult
from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator, TogetherGameStartGenerator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--games', type=int, default=65536)
    ap.add_argument('--ticks', type=int, default=200)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--opponent', choices=['self', 'random', 'staircase'], default='self',
                    help='player 2: the same policy (self-play), or a scripted bot that moves inside the tick kernel')
    args = ap.parse_args()
    dev = torch.device('cuda')
    torch.manual_seed(args.seed)

    dgen = EmptyDungeonGenerator(60, 10)
    game_state = TogetherGameStartGenerator(dgen).setup_game(args.games, seed=args.seed, device=dev)
    updater = BatchedUpdater(dgen, DungeonDespawningStrategy.Unreachable, max_ticks=500, auto_reset=True)

    policy = torch.nn.Sequential(torch.nn.Linear(_abi.OBS_LEN, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5)).to(dev)
    obs = torch.empty((args.games, 2, _abi.OBS_LEN), dtype=torch.int16, device=dev)
    result = torch.empty((args.games,), dtype=torch.uint8, device=dev)
    returns = torch.zeros((args.games, 2), device=dev)
    wins = torch.zeros(5, dtype=torch.int64, device=dev)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        updater.observe(game_state, stairs_radius=4, out=obs)                      # first view; ladder visible when near
        for _ in range(args.ticks):
            logits = policy(obs.float())                                           # [N, 2, 5]
            moves = (torch.distributions.Categorical(logits=logits).sample() + 1).to(torch.uint8)   # Move codes 1..5
            # one pass: the tick and what both players see of the new state (orx_step_observe); against a
            # scripted opponent its command is computed inside the same kernel (orx_step_bots)
            if args.opponent == 'self':
                updater.update_observe(game_state, moves.contiguous(), stairs_radius=4, out=result, obs_out=obs)
            else:
                kind = _abi.BOT_RANDOM if args.opponent == 'random' else _abi.BOT_STAIRCASE
                updater.update_with_bots(game_state, moves.contiguous(), _abi.BOT_NONE, kind, stairs_radius=4,
                                         out=result, obs_out=obs)
            r = result.long()
            returns[:, 0] += (r == UpdateResult.Player1Win).float() - (r == UpdateResult.Player2Win).float()
            returns[:, 1] -= (r == UpdateResult.Player1Win).float() - (r == UpdateResult.Player2Win).float()
            wins += torch.bincount(r, minlength=5)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    print(f'{args.games} games x {args.ticks} ticks in {el:.3f} s = {args.games * args.ticks / el:.3e} game-ticks/s '
          f'(policy included); finished: p1 {int(wins[2])}, p2 {int(wins[3])}, ties {int(wins[4])}; '
          f'episodes per lane: {float(game_state.episode.float().mean()):.2f}')


if __name__ == '__main__':
    main()
