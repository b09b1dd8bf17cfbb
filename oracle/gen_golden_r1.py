"""TEST INFRASTRUCTURE. Regression fixture for ruleset R1 (PARITY UNPINNED: no reference code exists for these rules, so this
is NOT a reference-derived golden vector): SHA-256 digests of trajectories of oracle/orx_r1_oracle.c, the plain-C restatement
of docs/RULESET_R1.md, committed as tests/golden/r1_oracle_digests.json so that an edit of the oracle (or of the spec it
restates) cannot pass unnoticed -- tests/test_r1_spec.py re-derives them on CPU, tests/test_gpu_r1.py from the CUDA kernels.
The digests of this file were first written by the round-1 ordering of the tick and re-derived unchanged after the round-2
reordering (pickups before descents, deaths before drops).

    python -m oracle.gen_golden_r1            # rewrites the fixture
"""
import hashlib
import json
import os

import numpy as np

CASES = [
    dict(name='default 60x10', n=600, ticks=400, cfg=dict(width=60, height=10, wall_density=26, seed=3, max_ticks=400, auto_reset=True), base=0),
    dict(name='small dense', n=600, ticks=300, cfg=dict(width=10, height=7, wall_density=40, seed=5, max_ticks=150, auto_reset=True), base=1 << 33),
    dict(name='open 7x7', n=600, ticks=300, cfg=dict(width=7, height=7, wall_density=0, seed=9, max_ticks=200, auto_reset=True), base=12345),
    dict(name='no reset', n=600, ticks=250, cfg=dict(width=24, height=8, wall_density=60, seed=11, max_ticks=120, auto_reset=False), base=7),
]
PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'r1_oracle_digests.json')


def commands(case_index, t, n):
    """Deterministic command bytes for tick t of a case (codes 0..7: the six commands and two invalid ones)."""
    return np.random.default_rng([case_index, t]).integers(0, 8, size=(n, 2), dtype=np.uint8)


def run_case(case_index, step, planes):
    """step(moves) -> results; planes() -> dict of numpy planes. Returns the hex digest of the whole trajectory."""
    from optimax_rogue_b200 import _abi
    case = CASES[case_index]
    h = hashlib.sha256()
    for t in range(case['ticks']):
        h.update(np.ascontiguousarray(step(commands(case_index, t, case['n']))).tobytes())
        if t % 20 == 0 or t == case['ticks'] - 1:
            p = planes()
            for name, _, _ in _abi.R1_PLANES:
                h.update(np.ascontiguousarray(p[name]).tobytes())
    return h.hexdigest()


def oracle_digests():
    from oracle import cport
    from optimax_rogue_b200 import _abi
    out = {}
    for k, case in enumerate(CASES):
        orc = cport.R1Oracle(case['n'], game_id_base=case['base'], **case['cfg'])
        orc.reset()
        out[case['name']] = run_case(k, orc.step, lambda: {name: getattr(orc.state, name) for name, _, _ in _abi.R1_PLANES})
    return out


if __name__ == '__main__':
    d = oracle_digests()
    with open(PATH, 'w') as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1))
