"""CPU oracle for the PRIME-RL per-token RL loss hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference (`pmahdavi/prime-rl-cse`) for the one hot path this repository
accelerates.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package
`prime_rl_cse_b200` never does (there is no CPU fallback).

Parity status: PINNED.  The reference is pure Python, importable in the build
container (`PYTHONPATH=/root/reference/src`).  `tests/golden/make_golden.py`
runs the *unmodified reference functions* on seeded inputs and commits their
outputs as fixtures under `tests/golden/`; `tests/test_oracle_golden.py` checks
this oracle against every one of them, plus the known-answer vectors of
SURVEY.md §8(c).  The reference's own tests pin shapes only
(`tests/unit/train/rl/test_loss.py:10-48`), so these fixtures are the pin.
"""

from .rl_loss_oracle import *  # noqa: F401,F403
