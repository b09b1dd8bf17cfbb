"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Runs the *live* reference (``/root/reference``: optimax_rogue.logic.updater.Updater,
optimax_rogue.logic.worldgen, optimax_rogue_bots) with its four random call
sites replaced by the shared Philox draw schedule of ``oracle/philox.py``, and
records per-tick traces. It exists only in the build container -- the GPU box
has no ``/root/reference`` -- so its outputs are committed as fixtures under
``tests/golden/`` by ``oracle/gen_golden.py``.

Nothing in the reference is modified on disk; the injection is done by
replacing module attributes (``random.shuffle``, ``random.choice``,
``numpy.random.randint``) and by wrapping two bound methods
(``dgen.spawn_dungeon`` to learn the depth, ``Updater.handle_descend`` to learn
the descender).
"""
import os
import random
import re
import sys
import types

import numpy as np

from . import philox as px

def _find_reference_root():
    """ORX_REFERENCE_ROOT, else the reference tree of the build container, else the copy oracle/make_ref.py
    ships to the GPU box (oracle/_ref, git-ignored)."""
    env = os.environ.get('ORX_REFERENCE_ROOT')
    if env:
        return env
    shipped = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
    for root in ('/root/reference', shipped):
        if os.path.isdir(os.path.join(root, 'optimax_rogue', 'logic')):
            return root
    return '/root/reference'


REFERENCE_ROOT = _find_reference_root()

# event codes shared with include/orx.h (ORX_EV_*)
EV_MOVE, EV_COMBAT, EV_DUNGEON, EV_DEATH, EV_DESCEND = 1, 2, 3, 4, 5


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'optimax_rogue', 'logic'))


def _install_inflection_shim():
    """networking/serializer.py:27 imports ``inflection`` (not installed) and
    uses only ``underscore`` (:95)."""
    if 'inflection' in sys.modules:
        return
    mod = types.ModuleType('inflection')

    def underscore(word):
        word = re.sub(r'([A-Z]+)([A-Z][a-z])', r'\1_\2', word)
        word = re.sub(r'([a-z\d])([A-Z])', r'\1_\2', word)
        return word.replace('-', '_').lower()
    mod.underscore = underscore
    sys.modules['inflection'] = mod


_REF = None


def load_reference():
    """Imports the reference packages (namespace packages, no install)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT}')
    _install_inflection_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import optimax_rogue.logic.updater as updater
    import optimax_rogue.logic.worldgen as worldgen
    import optimax_rogue.logic.updates as updates
    import optimax_rogue.logic.moves as moves
    import optimax_rogue.game.world as world
    import optimax_rogue.game.state as state
    import optimax_rogue.game.entities as entities
    import optimax_rogue_bots.randombot as randombot
    import optimax_rogue_bots.staircasebot as staircasebot
    import optimax_rogue_bots.bot as bot
    # the updater prints inside the hot path (updater.py:152,155,159,333; state.py:67-71)
    updater.print = lambda *a, **k: None
    state.print = lambda *a, **k: None
    _REF = types.SimpleNamespace(
        updater=updater, worldgen=worldgen, updates=updates, moves=moves, world=world,
        state=state, entities=entities, randombot=randombot, staircasebot=staircasebot, bot=bot)
    return _REF


class Injector:
    """Holds the per-game draw context and the replacement RNG functions."""

    def __init__(self, seed: int):
        self.seed = seed
        self.game_id = 0
        self.episode = 0
        self.tick = 0
        self.site = None          # ('reset',) | ('level', depth) | ('descend', p) | None
        self.q = 0                # running index inside the current site
        self.shuffle_calls = 0
        self.choice_calls = 0
        self._saved = None
        self.words_used = 0

    # -- replacement functions -------------------------------------------------------------
    def choice(self, seq):
        # G1: one bot draw per RandomBot per tick; harness order is p1 then p2
        w = px.block(self.seed, self.game_id, self.episode, px.DOM_TICK, px.SUB_TICK_MAIN,
                     self.tick)[self.choice_slot]
        self.words_used += 1
        return seq[px.bounded(w, len(seq))]

    def shuffle(self, x):
        # G2 (first call of the tick) / G3 (second call); CPython loop order
        first = self.shuffle_calls == 0
        self.shuffle_calls += 1
        q = 0
        for i in reversed(range(1, len(x))):
            if first:
                assert len(x) == 2
                w = px.block(self.seed, self.game_id, self.episode, px.DOM_TICK,
                             px.SUB_TICK_MAIN, self.tick)[2]
            else:
                w = px.seq_word(self.seed, self.game_id, self.episode, px.DOM_TICK,
                                px.SUB_NPC_SHUFFLE, self.tick, q)
                q += 1
            self.words_used += 1
            j = px.bounded(w, i + 1)
            x[i], x[j] = x[j], x[i]

    def randint(self, low, high=None):
        # G4 / G5
        if high is None:
            low, high = 0, low
        n = int(high) - int(low)
        site = self.site
        assert site is not None, 'np.random.randint called outside a known site'
        if site[0] == 'level':
            w = px.block(self.seed, self.game_id, self.episode, px.DOM_LEVEL, 0, site[1])[self.q]
            val = px.bounded(w, n)
        elif site[0] == 'reset':
            val = px.seq_bounded(self.seed, self.game_id, self.episode, px.DOM_RESET, 0, 0,
                                 self.q, n)
        elif site[0] == 'descend':
            val = px.seq_bounded(self.seed, self.game_id, self.episode, px.DOM_TICK,
                                 px.SUB_DESCEND + 64 * site[1], self.tick, self.q, n)
        else:
            raise AssertionError(site)
        self.q += 1
        self.words_used += 1
        return int(low) + val

    # -- patching ----------------------------------------------------------------------------
    def __enter__(self):
        self._saved = (random.shuffle, random.choice, np.random.randint)
        random.shuffle = self.shuffle
        random.choice = self.choice
        np.random.randint = self.randint
        return self

    def __exit__(self, *exc):
        random.shuffle, random.choice, np.random.randint = self._saved
        self._saved = None

    def wrap_dgen(self, dgen):
        orig = dgen.spawn_dungeon

        def spawn_dungeon(depth):
            saved = (self.site, self.q)
            self.site, self.q = ('level', depth), 0
            try:
                return orig(depth)
            finally:
                self.site, self.q = saved
        dgen.spawn_dungeon = spawn_dungeon
        return dgen

    def wrap_updater(self, upd, gs_ref):
        orig = upd.handle_descend

        def handle_descend(game_state, ent, result):
            p = 0 if ent.entity.iden == game_state.player_1_iden else 1
            saved = (self.site, self.q)
            self.site, self.q = ('descend', p), 0
            try:
                return orig(game_state, ent, result)
            finally:
                self.site, self.q = saved
        upd.handle_descend = handle_descend
        return upd


def make_fixed_generator(ref, tiles: np.ndarray):
    """A DungeonGenerator plugin (worldgen.py:9-26 interface) that returns the same
    tile grid at every depth. ``tiles`` is int[W,H] of Tile codes."""
    class FixedDungeonGenerator(ref.worldgen.DungeonGenerator):
        def __init__(self, tiles):
            super().__init__(tiles.shape[0], tiles.shape[1])
            self.tiles = np.asarray(tiles, dtype='int32')

        def spawn_dungeon(self, depth):
            return ref.world.Dungeon(self.tiles.copy())
    return FixedDungeonGenerator(tiles)


def make_flat_modifier(ref, parent, flat_damage=0, flat_armor=0, flat_max_health=0):
    """The smallest concrete ``Modifier`` (game/modifiers.py:92-150): only the three flat bonuses that
    ``Entity.on_tick`` folds into the attribles (game/attribles.py:21-43); its event hooks do nothing."""
    class FlatBonus(ref.entities.Modifier if hasattr(ref.entities, 'Modifier') else __import__(
            'optimax_rogue.game.modifiers', fromlist=['Modifier']).Modifier):
        name = 'flat bonus'
        description = 'test modifier: flat stat bonuses only'

        def copy(self, ent):
            return type(self)(ent, self.flat_armor, self.flat_max_health, self.flat_damage)

        def handles(self, event_name):
            return False

        def pre_event(self, event_name, game_state, args):
            return None

        def on_event(self, event_name, game_state, args, prevals):
            return None

        def post_event(self, event_name, game_state, args, prevals):
            return None
    return FlatBonus(parent, flat_armor, flat_max_health, flat_damage)


class ScriptedBot:
    """Plays a fixed list of move codes (then Stay)."""
    def __init__(self, ref, codes):
        self.ref = ref
        self.codes = list(codes)
        self.i = 0

    def move(self, game_state):
        code = self.codes[self.i] if self.i < len(self.codes) else 5
        self.i += 1
        # the reference's Move enum has codes 1..5 only (logic/moves.py:6-12); anything else is
        # not representable there, so scripted streams stay inside it
        return self.ref.moves.Move(code)


def _event_tuple(ref, ev):
    u = ref.updates
    if isinstance(ev, u.EntityPositionUpdate):
        code = EV_DESCEND if ev.depth != ev.old_depth else EV_MOVE
        if code == EV_DESCEND:
            assert ev.old_depth == ev.depth - 1
        return (code, ev.entity_iden, ev.posx, ev.posy, ev.depth)
    if isinstance(ev, u.EntityCombatUpdate):
        assert len(ev.tags) == 1 and not any(ev.attack_prevals) and not any(ev.defend_prevals)    # flat-bonus test modifiers return None
        return (EV_COMBAT, ev.attacker_iden, ev.defender_iden, int(next(iter(ev.tags))),
                ev.og_damage)
    if isinstance(ev, u.DungeonCreatedUpdate):
        sx, sy = ev.dungeon.staircase() if (ev.dungeon.tiles == 3).any() else (255, 255)
        return (EV_DUNGEON, 0, sx, sy, ev.depth)
    if isinstance(ev, u.EntityDeathUpdate):
        return (EV_DEATH, ev.entity_iden, 0, 0, 0)
    raise AssertionError(f'unexpected event {type(ev)}')


def _stairs_of(gs, depth):
    d = gs.world.dungeons[depth]
    hits = np.argwhere(d.tiles == 3)
    if len(hits) == 0:
        return (255, 255)
    return (int(hits[0][0]), int(hits[0][1]))


def snapshot(gs, result, events):
    """One trace record: the comparison unit of SURVEY.md 7/8(d) config 5."""
    p1, p2 = gs.player_1, gs.player_2
    rec = {
        'tick': gs.tick, 'result': int(result),
        'ent': [(p1.x, p1.y, p1.depth, p1.health), (p2.x, p2.y, p2.depth, p2.health)],
        'stairs': [_stairs_of(gs, p1.depth), _stairs_of(gs, p2.depth)],
        'events': list(events),
        'npcs': [(e.iden, e.depth, e.x, e.y, e.health) for e in gs.entities
                 if e.iden not in (gs.player_1_iden, gs.player_2_iden)],
    }
    # what each player is shown: GameState.view_for (game/state.py:53-58) -- the entities on the viewer's depth, that
    # level alone, the tick. Not part of the digests; pins the observation records (orx_observe, orx_observe_npc).
    views = []
    for viewer in (p1, p2):
        v = gs.view_for(viewer)
        views.append({'tick': v.tick, 'levels': sorted(v.world.dungeons.keys()), 'stairs': _stairs_of(v, viewer.depth),
                      'ents': sorted((e.iden, e.depth, e.x, e.y, e.health) for e in v.entities)})
    rec['views'] = views
    return rec


def play_episode(seed, game_id, episode=0, *, bots=('random', 'random'), width=60, height=10,
                 start='together', p_depths=(0, 1000), despawn='unreachable', max_ticks=512,
                 fixed_tiles=None, npcs=(), scripts=None, limit_ticks=None,
                 hp=10, damage=2, armor=1, want_order=False, place=None, flat=None):
    """Plays one episode on the live reference under injected draws.

    Returns the list of records: record 0 is the post-reset state, record t>0 the
    state after the t-th ``Updater.update``.
    """
    ref = load_reference()
    inj = Injector(seed)
    inj.game_id, inj.episode = game_id, episode
    with inj:
        if fixed_tiles is not None:
            dgen = make_fixed_generator(ref, fixed_tiles)
        else:
            dgen = ref.worldgen.EmptyDungeonGenerator(width, height)
        inj.wrap_dgen(dgen)
        if start == 'together':
            gen = ref.worldgen.TogetherGameStartGenerator(dgen)
        else:
            gen = ref.worldgen.SeparatedGameStartGenerator(dgen, p_depths[0], p_depths[1])
        inj.site, inj.q = ('reset',), 0
        gs = gen.setup_game()
        inj.site = None
        hp2, dmg2, arm2 = (v if isinstance(v, (tuple, list)) else (v, v) for v in (hp, damage, armor))
        for k, ent in enumerate((gs.player_1, gs.player_2)):
            ent.health, ent.base_max_health = hp2[k], hp2[k]
            ent.base_damage, ent.base_armor = dmg2[k], arm2[k]
        if flat is not None:       # ((flat_damage, flat_armor, flat_max_health) of p1, the same of p2): one modifier each
            for ent, (fd, fa, fm) in zip((gs.player_1, gs.player_2), flat):
                ent.modifiers.append(make_flat_modifier(ref, ent, fd, fa, fm))
        if place is not None:
            for ent, (px_, py_) in zip((gs.player_1, gs.player_2), place):
                ent.x, ent.y = int(px_), int(py_)
            gs.pos_lookup = dict(((e.depth, e.x, e.y), e) for e in gs.entities)
        for k, (nd, nx, ny, nhp) in enumerate(npcs):
            gs.add_entity(ref.entities.Entity(3 + k, nd, nx, ny, nhp, nhp, 0, 0, [], dict()))
        strat = (ref.updater.DungeonDespawningStrategy.Unreachable if despawn == 'unreachable'
                 else ref.updater.DungeonDespawningStrategy.Unused)
        upd = ref.updater.Updater(dgen, strat, max_ticks)
        inj.wrap_updater(upd, gs)

        def mk(kind, iden, idx):
            if kind == 'random':
                return ref.randombot.RandomBot(iden)
            if kind == 'staircase':
                return ref.staircasebot.StaircaseBot(iden)
            if kind == 'script':
                return ScriptedBot(ref, scripts[idx])
            raise ValueError(kind)
        b1, b2 = mk(bots[0], 1, 0), mk(bots[1], 2, 1)

        trace = [snapshot(gs, ref.updater.UpdateResult.InProgress, [])]
        n = 0
        moves_log = []
        while True:
            inj.tick = gs.tick
            inj.shuffle_calls = 0
            gs.on_tick()                         # server/main.py:111
            inj.choice_slot = 0
            m1 = b1.move(gs)
            inj.choice_slot = 1
            m2 = b2.move(gs)
            moves_log.append((int(m1), int(m2)))
            res, evs = upd.update(gs, m1, m2)    # networking/server.py:126
            evt = [_event_tuple(ref, e) for e in evs]
            if want_order:
                assert [e.order for e in evs] == list(range(upd.current_update_order - len(evs),
                                                            upd.current_update_order))
            trace.append(snapshot(gs, res, evt))
            n += 1
            if res != ref.updater.UpdateResult.InProgress:
                break
            if limit_ticks is not None and n >= limit_ticks:
                break
    return trace, moves_log


# ---- digests --------------------------------------------------------------------------------
FNV_OFFSET = 0xcbf29ce484222325
FNV_PRIME = 0x100000001b3
M64 = (1 << 64) - 1


def record_values(rec, max_events=4):
    """The flat integer list a record contributes to the digest (fixed layout)."""
    vals = [rec['tick'], rec['result']]
    for e in rec['ent']:
        vals.extend(e)
    for s in rec['stairs']:
        vals.extend(s)
    vals.append(len(rec['events']))
    for k in range(max_events):
        if k < len(rec['events']):
            vals.extend(rec['events'][k])
        else:
            vals.extend((0, 0, 0, 0, 0))
    return vals


def digest(trace, with_events=True, max_events=4):
    """FNV-1a style fold (one 32-bit value per step) over all records of an episode."""
    h = FNV_OFFSET
    for rec in trace:
        vals = record_values(rec, max_events)
        if not with_events:
            vals = vals[:14]
        for v in vals:
            h = ((h ^ (v & 0xffffffff)) * FNV_PRIME) & M64
    return h
