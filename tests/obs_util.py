"""numpy restatements of the observation records (include/orx.h: orx_observe, orx_observe_npc), shared by the GPU tests
(CUDA == restatement) and tests/test_observation_vs_reference.py (restatement == the live reference's
GameState.view_for, game/state.py:53-58): the two together pin the CUDA observations to the reference."""
import numpy as np


def expected_obs(p, radius):
    """planes (dict of numpy arrays as BatchedGameState.planes_cpu returns them) -> int16[n, 2, 12]."""
    n = p['pos'].shape[0]
    obs = np.zeros((n, 2, 12), dtype=np.int64)
    same = p['depth'][:, 0] == p['depth'][:, 1]
    for pl in range(2):
        o = 1 - pl
        x, y = p['pos'][:, 2 * pl].astype(int), p['pos'][:, 2 * pl + 1].astype(int)
        sx, sy = p['stairs'][:, 2 * pl].astype(int), p['stairs'][:, 2 * pl + 1].astype(int)
        vis = (sx != 255) & ((radius < 0) | (np.maximum(np.abs(sx - x), np.abs(sy - y)) <= radius))
        cols = [x, y, np.minimum(p['depth'][:, pl], 32767), p['hp'][:, pl], same,
                np.where(same, p['pos'][:, 2 * o].astype(int), -1), np.where(same, p['pos'][:, 2 * o + 1].astype(int), -1),
                np.where(same, p['hp'][:, o], 0), vis, np.where(vis, sx, -1), np.where(vis, sy, -1),
                np.minimum(p['tick'], 32767)]
        for c, v in enumerate(cols):
            obs[:, pl, c] = v
    return obs.astype(np.int16)


def expected_npc_obs(p):
    """planes -> int16[n, 2, n_npc, 4] = on_my_depth, x, y, health per player and NPC slot."""
    n, e = p['npc_depth'].shape
    want = np.zeros((n, 2, e, 4), np.int16)
    for pl in range(2):
        here = (p['npc_depth'] >= 0) & (p['npc_depth'] == p['depth'][:, pl:pl + 1])
        want[:, pl, :, 0] = here
        want[:, pl, :, 1] = np.where(here, p['npc_pos'][:, :, 0].astype(int), -1)
        want[:, pl, :, 2] = np.where(here, p['npc_pos'][:, :, 1].astype(int), -1)
        want[:, pl, :, 3] = np.where(here, p['npc_hp'], 0)
    return want
