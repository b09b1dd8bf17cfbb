"""Shared helpers for parity tests: run an engine (oracle or CUDA) lane-by-lane into the
record format of oracle/ref_harness.snapshot, and the vectorised digest."""
import numpy as np

from oracle import cport

BOT_CODES = {'none': 0, 'random': 1, 'staircase': 2}


def records_from_planes(pos, hp, depth, stairs, tick, result, events, lane):
    """One ref_harness-style record for `lane` from SoA planes (numpy)."""
    evs = []
    if events is not None:
        dec = cport.decode_events(events[lane])
        for k in range(dec.shape[0]):
            if dec[k, 0] == 0:
                break
            evs.append(tuple(int(v) for v in dec[k]))
    return {
        'tick': int(tick[lane]), 'result': int(result[lane]),
        'ent': [(int(pos[lane, 0]), int(pos[lane, 1]), int(depth[lane, 0]), int(hp[lane, 0])),
                (int(pos[lane, 2]), int(pos[lane, 3]), int(depth[lane, 1]), int(hp[lane, 1]))],
        'stairs': [(int(stairs[lane, 0]), int(stairs[lane, 1])),
                   (int(stairs[lane, 2]), int(stairs[lane, 3]))],
        'events': evs,
    }


def strip(rec):
    return {k: rec[k] for k in ('tick', 'result', 'ent', 'stairs', 'events')}


def oracle_episode(cfg, game_id, bots=('random', 'random'), scripts=None, limit_ticks=None,
                   npcs=(), place=None, flat=None):
    """Plays one game on the C oracle; returns (trace, moves_log) like rh.play_episode."""
    orc = cport.Oracle(cfg, 1, game_id_base=game_id)
    orc.reset()
    s = orc.state
    if flat is not None:
        s.enable_flat_bonuses()[0] = flat
    if place is not None:
        s.pos[0] = (place[0][0], place[0][1], place[1][0], place[1][1])
    for k, (nd, nx, ny, nhp) in enumerate(npcs):
        s.npc_depth[0, k] = nd
        s.npc_pos[0, k] = (nx, ny)
        s.npc_hp[0, k] = nhp
    trace = [records_from_planes(s.pos, s.hp, s.depth, s.stairs, s.tick, s.status, None, 0)]
    moves_log = []
    t = 0
    while True:
        if scripts is not None:
            mv = np.array([[scripts[p][t] if t < len(scripts[p]) else 5 for p in range(2)]], np.uint8)
            for p in range(2):
                if bots[p] != 'script':
                    one = orc.bot_moves(BOT_CODES[bots[0]] if p == 0 else 0,
                                        BOT_CODES[bots[1]] if p == 1 else 0)
                    mv[0, p] = one[0, p]
        else:
            mv = orc.bot_moves(BOT_CODES[bots[0]], BOT_CODES[bots[1]])
        moves_log.append((int(mv[0, 0]), int(mv[0, 1])))
        res, ev = orc.step(mv, want_events=True)
        trace.append(records_from_planes(s.pos, s.hp, s.depth, s.stairs, s.tick, res, ev, 0))
        t += 1
        if res[0] != 1:
            break
        if limit_ticks is not None and t >= limit_ticks:
            break
    return trace, moves_log


FNV_OFFSET = np.uint64(0xcbf29ce484222325)
FNV_PRIME = np.uint64(0x100000001b3)


class BatchDigest:
    """Vectorised twin of rh.digest: folds one record per lane per call."""

    def __init__(self, n, with_events=True, max_events=4):
        self.h = np.full(n, FNV_OFFSET, np.uint64)
        self.with_events = with_events
        self.max_events = max_events

    def _fold(self, v, active):
        v = (np.asarray(v).astype(np.int64) & 0xffffffff).astype(np.uint64)
        with np.errstate(over='ignore'):
            nh = (self.h ^ v) * FNV_PRIME
        self.h = np.where(active, nh, self.h)

    def update(self, pos, hp, depth, stairs, tick, result, events, active):
        f = self._fold
        f(tick, active); f(result, active)
        for p in range(2):
            f(pos[:, 2 * p], active); f(pos[:, 2 * p + 1], active)
            f(depth[:, p], active); f(hp[:, p], active)
        for k in range(4):
            f(stairs[:, k], active)
        if not self.with_events:
            return
        if events is None:
            f(np.zeros_like(tick), active)
            for _ in range(self.max_events * 5):
                f(np.zeros_like(tick), active)
            return
        dec = cport.decode_events(events)          # [n, E, 5]
        nev = (dec[:, :, 0] != 0).sum(axis=1)
        f(nev, active)
        for k in range(self.max_events):
            for c in range(5):
                f(dec[:, k, c], active)
