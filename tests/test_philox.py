"""Philox4x32-10 known-answer vectors (Random123 reference outputs) for the
pure-Python generator the reference harness uses and for the C oracle, plus the schedule's
bounded() map."""
import numpy as np

from oracle import cport
from oracle import philox as px

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_python_philox_kat():
    for ctr, key, out in KATS:
        assert px.philox4x32_10(ctr, key) == out


def test_c_oracle_philox_kat():
    for ctr, key, out in KATS:
        assert cport.philox(ctr, key) == out


def test_python_and_c_agree_on_random_counters():
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr = tuple(int(x) for x in rng.integers(0, 1 << 32, 4))
        key = tuple(int(x) for x in rng.integers(0, 1 << 32, 2))
        assert px.philox4x32_10(ctr, key) == cport.philox(ctr, key)


def test_bounded_is_mulhi():
    assert px.bounded(0, 5) == 0
    assert px.bounded(0xffffffff, 5) == 4
    assert px.bounded(0x80000000, 2) == 1
    assert px.bounded(0x7fffffff, 2) == 0
    assert px.bounded(0x33333333, 5) == 0 and px.bounded(0x33333334, 5) == 1


def test_block_counter_layout():
    # game id high bits, sub-block and domain share counter word 1
    gid = (0x2ABCDE << 32) | 0x12345678
    got = px.block(0x1122334455667788, gid, 9, px.DOM_RESET, 3, 77)
    c1 = 0x2ABCDE | (3 << 22) | (px.DOM_RESET << 30)
    assert got == px.philox4x32_10((0x12345678, c1, 9, 77), (0x55667788, 0x11223344))
