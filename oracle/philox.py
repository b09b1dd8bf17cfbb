"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Philox4x32-10 counter-based generator and the shared *draw schedule* that both
the reference (through injection, see ``oracle/ref_harness.py``) and the CUDA
kernels (``optimax_rogue_b200/csrc/orx_rng.cuh``) consume.

The reference draws from CPython ``random`` and ``numpy.random`` at four call
sites:

  G1  optimax_rogue_bots/randombot.py:21   random.choice(self.moves)
  G2  optimax_rogue/logic/updater.py:114   random.shuffle(updents)
  G3  optimax_rogue/logic/updater.py:127   random.shuffle(npcs)
  G4  optimax_rogue/logic/worldgen.py:39-40  np.random.randint(1, W-2 / H-2)
  G5  optimax_rogue/game/world.py:62       np.random.randint(#Ground)

Every one of those calls is replaced by ``bounded(word, n) = (word * n) >> 32``
over exactly one 32-bit word of the Philox stream, selected by the schedule
below, so the number of words consumed per reference call is static.

Counter layout (key = (seed_lo, seed_hi)):

  c0 = game_id & 0xffffffff
  c1 = (game_id >> 32) & 0x3fffff | sub << 22 | domain << 30      (game_id < 2**54)
  c2 = episode
  c3 = index   (tick for TICK, depth for LEVEL, 0 for RESET)

Domains / sub-blocks (each block yields four words w0..w3):

  TICK  (0), sub 0            w0 = p1 RandomBot, w1 = p2 RandomBot, w2 = initiative, w3 reserved
  TICK  (0), sub 1..7         NPC shuffle draw q -> word q&3 of sub 1 + (q>>2)
  TICK  (0), sub 64*(1+p)+k   descend-spawn try r of player p (0/1): word r&3 of k = r>>2  (r < 256)
  LEVEL (1), sub 0            w0 = stair x, w1 = stair y
  RESET (2), sub k            spawn draw q: word q&3 of k = q>>2 (q = 0: p1, q = 1 + r: p2 try r; q < 256)

Tries beyond 255 (probability ~ 463**-255) fall back to ``(try_255 + (r-255)) % n``.
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

DOM_TICK = 0
DOM_LEVEL = 1
DOM_RESET = 2

SUB_TICK_MAIN = 0
SUB_NPC_SHUFFLE = 1
SUB_DESCEND = 64  # + 64 * player_index + block
MAX_TRIES = 256


def philox4x32_10(ctr, key):
    """ctr = (c0, c1, c2, c3), key = (k0, k1); returns 4 words."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return (c0, c1, c2, c3)


def bounded(word, n):
    """Maps one 32-bit word to [0, n) without rejection."""
    return (word * n) >> 32


def block(seed, game_id, episode, domain, sub, index):
    """The four words of one schedule block."""
    assert 0 <= game_id < (1 << 54)
    c0 = game_id & MASK
    c1 = ((game_id >> 32) & 0x3FFFFF) | ((sub & 0xFF) << 22) | ((domain & 3) << 30)
    return philox4x32_10((c0, c1, episode & MASK, index & MASK),
                         (seed & MASK, (seed >> 32) & MASK))


def seq_word(seed, game_id, episode, domain, sub_base, index, q):
    """Word q of a draw sequence that starts at block ``sub_base``."""
    if q < MAX_TRIES:
        return block(seed, game_id, episode, domain, sub_base + (q >> 2), index)[q & 3]
    raise OverflowError  # callers handle the fallback on the bounded value


def seq_bounded(seed, game_id, episode, domain, sub_base, index, q, n):
    """bounded() of draw q in a sequence, with the deterministic fallback."""
    if q < MAX_TRIES:
        return bounded(seq_word(seed, game_id, episode, domain, sub_base, index, q), n)
    last = bounded(seq_word(seed, game_id, episode, domain, sub_base, index, MAX_TRIES - 1), n)
    return (last + (q - (MAX_TRIES - 1))) % n
