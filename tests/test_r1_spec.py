"""Ruleset R1 scenarios: pins oracle/orx_r1_oracle.c to the written spec (docs/RULESET_R1.md) with
hand-computed outcomes for each README rule (readme.md:44-48,69-74). R1 has no reference
implementation -- parity with the reference is unpinned by construction."""
import numpy as np

from oracle import cport

UP, RIGHT, DOWN, LEFT, STAY, HEAL = 1, 2, 3, 4, 5, 6


def game(**kw):
    cfg = dict(width=12, height=8, wall_density=0, seed=1)
    cfg.update(kw)
    o = cport.R1Oracle(1, **cfg)
    o.reset()
    s = o.state
    s.lvl_stairs[0] = 10 | (6 << 8) | (10 << 16) | (6 << 24)      # staircase out of the way
    return o, s


def quiet_seed(ticks):
    """A seed whose first `ticks` ticks spawn no enemy (the spawn draw depends only on seed, game,
    episode and tick), so that a scenario sees only the entities it placed."""
    for seed in range(1000):
        o, s = game(seed=seed)
        put(s, 0, 2, 2); put(s, 1, 9, 6)
        for _ in range(ticks):
            o.step([[STAY, STAY]])
        if ((s.ent_loc[0, 2:10] >> 16) & 1).sum() == 0:
            return seed
    raise AssertionError('no quiet seed')


def put(s, lane, x, y, depth=0, hp=10, aux=9, kind=0):
    s.ent_loc[0, lane] = x | (y << 8) | (1 << 16) | ((kind & 3) << 17 if lane >= 10 else 0)
    s.ent_depth[0, lane] = depth
    s.ent_stat[0, lane] = (hp & 0xFFFF) | ((aux << 16) if lane < 10 else 0)


def hp(s, lane):
    return int(np.int16(int(s.ent_stat[0, lane]) & 0xFFFF))


def mana(s, lane):
    return int(s.ent_stat[0, lane]) >> 16


def xy(s, lane):
    return int(s.ent_loc[0, lane]) & 255, (int(s.ent_loc[0, lane]) >> 8) & 255


def cd(s, p):
    return (int(s.pl_b[0, p]) >> 24) & 255


def test_attack_on_a_staying_defender_is_negated_and_stuns_the_attacker():
    o, s = game()
    put(s, 0, 3, 3); put(s, 1, 4, 3)
    o.step([[RIGHT, STAY]])
    assert hp(s, 1) == 10 and xy(s, 0) == (3, 3)
    assert cd(s, 0) == 1 and mana(s, 0) == 6           # a third of the mana bar was committed
    # next turn: the stunned attacker can neither attack (measured as Stay) nor defend
    o.step([[RIGHT, LEFT]])
    assert hp(s, 1) == 10
    assert hp(s, 0) == 10 - (2 + 3 - 1)                # full damage: 2 + min(9, 9/3) mana - armor 1
    assert cd(s, 0) == 0 and cd(s, 1) == 0


def test_mutual_attack_is_half_damage_and_a_three_turn_cooldown():
    o, s = game()
    put(s, 0, 3, 3); put(s, 1, 4, 3)
    o.step([[RIGHT, LEFT]])
    assert hp(s, 0) == 8 and hp(s, 1) == 8             # (2 + 3 - 1) // 2
    assert cd(s, 0) == 3 and cd(s, 1) == 3 and mana(s, 0) == 6 and mana(s, 1) == 6
    assert xy(s, 0) == (3, 3) and xy(s, 1) == (4, 3)
    o.step([[STAY, STAY]])
    assert cd(s, 0) == 2


def test_attacking_a_tile_the_defender_leaves_deals_nothing():
    o, s = game()
    put(s, 0, 3, 3); put(s, 1, 4, 3)
    o.step([[RIGHT, UP]])
    assert hp(s, 1) == 10 and xy(s, 1) == (4, 2) and xy(s, 0) == (3, 3) and mana(s, 0) == 9


def test_attacking_the_defenders_new_location_is_full_damage():
    o, s = game()
    put(s, 0, 3, 3); put(s, 1, 5, 3)
    o.step([[RIGHT, LEFT]])                            # both step onto (4, 3)
    assert hp(s, 0) == 6 and hp(s, 1) == 6
    assert xy(s, 0) == (3, 3) and xy(s, 1) == (5, 3)
    assert cd(s, 0) == 0 and cd(s, 1) == 0


def test_heal_spends_up_to_a_third_of_the_mana_bar():
    o, s = game()
    put(s, 0, 3, 3, hp=5); put(s, 1, 8, 5, hp=9, aux=2)
    o.step([[HEAL, HEAL]])
    assert hp(s, 0) == 8 and mana(s, 0) == 6
    assert hp(s, 1) == 10 and mana(s, 1) == 0          # health is capped, the committed mana is spent


def test_mana_regenerates_every_fourth_tick():
    o, s = game()
    put(s, 0, 3, 3, aux=0); put(s, 1, 8, 5, aux=0)
    got = []
    for _ in range(8):
        o.step([[STAY, STAY]])
        got.append(mana(s, 0))
    assert got == [0, 0, 0, 1, 1, 1, 1, 2]             # ticks 1..8: +1 when tick % 4 == 0


def test_separation_damage_grows_linearly_for_the_player_behind():
    o, s = game()
    put(s, 0, 3, 3, depth=0); put(s, 1, 8, 5, depth=2)
    s.sep[0] = 31
    o.step([[STAY, STAY]])
    assert int(s.sep[0]) == 32 and hp(s, 0) == 10 - 32 // 16 and hp(s, 1) == 10
    s.ent_depth[0, 1] = 0
    o.step([[STAY, STAY]])
    assert int(s.sep[0]) == 0


def test_item_pickup_gives_a_flat_bonus_and_slots_are_finite():
    o, s = game()
    put(s, 0, 3, 3); put(s, 1, 8, 5)
    put(s, 10, 4, 3, kind=0); put(s, 11, 5, 3, kind=2)
    o.step([[RIGHT, STAY]])
    assert int(s.pl_c[0, 0]) & 255 == 3 and (int(s.pl_b[0, 0]) >> 16) & 255 == 1 and int(s.ent_loc[0, 10]) == 0
    o.step([[RIGHT, STAY]])
    assert int(np.int16(int(s.pl_a[0, 0]) & 0xFFFF)) == 12 and hp(s, 0) == 12
    s.pl_b[0, 0] = (int(s.pl_b[0, 0]) & ~(255 << 16)) | (4 << 16)     # slots full
    put(s, 12, 6, 3, kind=1)
    o.step([[RIGHT, STAY]])
    assert (int(s.ent_loc[0, 12]) >> 16) & 1 == 1 and (int(s.pl_c[0, 0]) >> 8) & 255 == 1


def test_enemy_chases_attacks_dies_and_three_kills_level_up():
    o, s = game(seed=quiet_seed(6))
    put(s, 0, 3, 3); put(s, 1, 9, 6)
    put(s, 2, 6, 3, hp=1, aux=0)
    o.step([[STAY, STAY]])
    assert xy(s, 2) == (5, 3)                          # steps along the larger axis towards player 1
    o.step([[STAY, STAY]])
    assert xy(s, 2) == (4, 3)
    o.step([[STAY, STAY]])
    assert hp(s, 0) == 10                              # a staying player negates the enemy's attack
    s.pl_b[0, 0] = (int(s.pl_b[0, 0]) & 0x00FFFFFF) | (2 << 24)       # on cooldown: cannot defend
    o.step([[STAY, STAY]])
    assert hp(s, 0) == 9                               # enemy damage 2 + depth/4 minus armor 1
    s.pl_b[0, 0] = (int(s.pl_b[0, 0]) & 0x00FFFF00) | 2               # cooldown off, xp = 2
    put(s, 0, 3, 3, hp=4, aux=1)
    o.step([[RIGHT, STAY]])                            # enemies never negate: full damage kills it
    assert (int(s.ent_loc[0, 2]) >> 16) & 1 == 0
    assert int(s.pl_b[0, 0]) & 255 == 0 and (int(s.pl_b[0, 0]) >> 8) & 255 == 2
    assert hp(s, 0) == 10 and mana(s, 0) == 9          # levelling refills health and mana


def test_double_death_draws_the_winner():
    wins = set()
    for seed in range(12):
        o, s = game(seed=seed)
        put(s, 0, 3, 3, hp=2); put(s, 1, 5, 3, hp=2)
        res = o.step([[RIGHT, LEFT]])
        assert res[0] in (2, 3)
        wins.add(int(res[0]))
    assert wins == {2, 3}


def test_walls_are_a_pure_function_of_the_level_key():
    a = cport.R1Oracle(64, width=60, height=10, wall_density=64, seed=1)
    b = cport.R1Oracle(64, width=60, height=10, wall_density=0, seed=1)
    a.reset(); b.reset()
    assert np.array_equal(a.state.lvl_key, b.state.lvl_key)           # the key does not depend on the density
    assert np.array_equal(a.state.lvl_stairs, b.state.lvl_stairs)
    assert not np.array_equal(a.state.ent_loc, b.state.ent_loc)       # but spawn tiles avoid the walls
    # a dense map still plays: nobody is ever placed on a wall, games keep a legal status
    big = cport.R1Oracle(500, width=20, height=8, wall_density=100, seed=5, max_ticks=80, auto_reset=True)
    big.reset()
    st = big.rollout(200)
    assert int(st[0]) == 500 * 200 and (big.state.status == 1).all()


def test_observation_layout_and_ladder_visibility():
    o, s = game(wall_density=0)
    put(s, 0, 3, 3); put(s, 1, 8, 5, hp=7)
    put(s, 2, 4, 4, hp=2, aux=0); put(s, 10, 2, 2, kind=1)
    s.lvl_stairs[0] = 6 | (3 << 8) | (6 << 16) | (3 << 24)
    obs = o.observe(radius=3)
    a, b = obs[0, 0], obs[0, 1]
    assert a[:5].tolist() == [3, 3, 0, 10, 9] and a[16:20].tolist() == [1, 8, 5, 7]
    assert a[20:23].tolist() == [1, 6, 3]                 # staircase 3 tiles away: visible (readme.md:44)
    assert b[20:23].tolist() == [1, 6, 3]                 # Chebyshev distance max(2, 2) = 2
    assert o.observe(radius=1)[0, 0, 20:23].tolist() == [0, -1, -1]
    assert a[23:26].tolist() == [4, 4, 2] and a[26:29].tolist() == [-1, -1, 0]
    assert a[47:50].tolist() == [2, 2, 1]
    # 7x7 wall window of player 1 at (3, 3): column x = 0 is the border wall => bit 0 of every row
    bits = int(a[59]) & 0xFFFF | (int(a[60]) & 0xFFFF) << 16 | (int(a[61]) & 0xFFFF) << 32 | (int(a[62]) & 0xFFFF) << 48
    for row in range(7):
        y = 3 + row - 3
        for col in range(7):
            x = 3 + col - 3
            wall = x <= 0 or y <= 0 or x >= 11 or y >= 7
            assert (bits >> (row * 7 + col)) & 1 == int(wall), (row, col)


# ---------------------------------------------------------------- replication log (include/orx.h, orx_r1_step_events)
def _shadow_from_state(s, g):
    """iden -> [x, y, depth, hp-or-kind] of every living lane of game g, plus the players' max health."""
    sh = {}
    for lane in range(14):
        loc = int(s.ent_loc[g, lane])
        if (loc >> 16) & 1:
            val = int(np.int16(int(s.ent_stat[g, lane]) & 0xFFFF)) if lane < 10 else (loc >> 17) & 3
            sh[lane + 1] = [loc & 255, (loc >> 8) & 255, int(s.ent_depth[g, lane]), val]
    max_hp = [int(np.int16(int(s.pl_a[g, p]) & 0xFFFF)) for p in range(2)]
    return sh, max_hp


def test_replication_log_reproduces_the_state():
    """A replica that starts from the same state and applies only the tick's GameStateUpdate records (the way the
    reference's clients do, networking/... -> update.apply, logic/updates.py) ends up with the same entities, tiles,
    depths and health as the oracle -- every tick, over bot-driven and random commands, descents, drops, pickups,
    level-ups, vanishing levels and separation damage. Ticks that end a game re-synchronise (auto-reset re-deals it)."""
    from optimax_rogue_b200 import _abi
    from optimax_rogue_b200.logic import updates as U
    n = 48
    o = cport.R1Oracle(n, width=14, height=9, wall_density=20, seed=11, max_ticks=300, auto_reset=True)
    o.reset()
    s = o.state
    rng = np.random.default_rng(2)
    shadows = [_shadow_from_state(s, g) for g in range(n)]
    seen = set()
    orders = [0] * n
    for t in range(700):
        mv = rng.integers(1, 7, size=(n, 2), dtype=np.uint8)
        if t % 4:
            o.bot_moves(_abi.BOT_STAIRCASE, _abi.BOT_RANDOM if t % 8 < 4 else _abi.BOT_STAIRCASE, mv)
        res, ev = o.step_events(mv)
        rec = U.unpack_events(ev)
        for g in range(n):
            ups = U.decode_r1_events(rec[g], first_order=orders[g])
            assert [u.order for u in ups] == list(range(orders[g], orders[g] + len(ups)))
            orders[g] += len(ups)
            sh, max_hp = shadows[g]
            for u in ups:
                seen.add(type(u).__name__)
                if isinstance(u, U.EntityPositionUpdate):
                    sh[u.entity_iden][:3] = [u.posx, u.posy, u.depth]
                elif isinstance(u, U.EntityCombatUpdate):
                    sh[u.defender_iden][3] -= u.og_damage
                elif isinstance(u, U.EntityHealthUpdate):
                    sh[u.entity_iden][3] += u.amount
                elif isinstance(u, U.EntitySpawnUpdate):
                    e = u.entity
                    assert e.iden not in sh
                    sh[e.iden] = [e.x, e.y, e.depth, e.item_kind if e.iden > 10 else e.health]
                elif isinstance(u, U.EntityDeathUpdate):
                    del sh[u.entity_iden]
                elif isinstance(u, U.EntityModifierAddedUpdate):
                    del sh[u.item_iden]                        # the item left the ground
                    max_hp[u.entity_iden - 1] += u.modifier.flat_max_health
                elif isinstance(u, U.EntityEventUpdate):
                    if u.args['levels_gained']:
                        sh[u.entity_iden][3] = max_hp[u.entity_iden - 1]
            if res[g] != 1:                                     # the game ended: players' DEATH records removed them; re-deal
                shadows[g] = _shadow_from_state(s, g)
                continue
            want, want_max = _shadow_from_state(s, g)
            assert sh == want, (t, g, sh, want)
            assert max_hp == want_max
    assert seen == {'EntityPositionUpdate', 'EntityCombatUpdate', 'EntityHealthUpdate', 'EntitySpawnUpdate', 'EntityDeathUpdate',
                    'EntityModifierAddedUpdate', 'EntityEventUpdate', 'DungeonCreatedUpdate'}


def test_event_capacity_drops_but_never_corrupts():
    """With fewer slots than records the first max_events records are the same and nothing is written past them."""
    o1 = cport.R1Oracle(32, width=14, height=9, wall_density=20, seed=3, max_ticks=200, auto_reset=True)
    o2 = cport.R1Oracle(32, width=14, height=9, wall_density=20, seed=3, max_ticks=200, auto_reset=True)
    o1.reset(); o2.reset()
    rng = np.random.default_rng(5)
    for _ in range(200):
        mv = rng.integers(1, 7, size=(32, 2), dtype=np.uint8)
        r1, e1 = o1.step_events(mv)
        r2, e2 = o2.step_events(mv, max_events=3)
        assert np.array_equal(r1, r2) and np.array_equal(e1[:, :3], e2)
    for name in ('ent_loc', 'ent_stat', 'ent_depth'):
        assert np.array_equal(getattr(o1.state, name), getattr(o2.state, name))


def test_r1_bots():
    """RandomBot = the fused rollout's policy (same draws), StaircaseBot walks to the staircase of its level."""
    o = cport.R1Oracle(64, seed=9, max_ticks=0, auto_reset=True)
    o.reset()
    ref = cport.R1Oracle(64, seed=9, max_ticks=0, auto_reset=True)
    ref.reset()
    for _ in range(50):
        o.step(o.bot_moves(1, 1))
    ref.rollout(50)
    for name in ('ent_loc', 'ent_stat', 'ent_depth', 'tick', 'episode'):
        assert np.array_equal(getattr(o.state, name), getattr(ref.state, name)), name
    o2, s2 = game()
    put(s2, 0, 2, 2); put(s2, 1, 9, 6)
    s2.lvl_stairs[0] = 7 | (3 << 8) | (9 << 16) | (2 << 24)
    mv = o2.bot_moves(2, 2)
    assert mv.tolist() == [[RIGHT, UP]]
    keep = np.full((1, 2), 77, np.uint8)
    assert o2.bot_moves(0, 2, keep).tolist() == [[77, UP]]


def test_oracle_reproduces_its_committed_trajectory_digests():
    """Regression fixture (tests/golden/r1_oracle_digests.json, oracle/gen_golden_r1.py): an edit of the R1 oracle or of the
    spec it restates must not pass unnoticed. Not a reference-derived vector -- R1 has no reference code (parity unpinned)."""
    import json
    from oracle import gen_golden_r1 as gg
    assert gg.oracle_digests() == json.load(open(gg.PATH))
