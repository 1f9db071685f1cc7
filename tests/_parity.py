"""Parity bookkeeping for the GPU tests: every whole-network comparison goes through ``within`` so that the measured
error is printed (``pytest -s`` / ``-rP``) next to its bound and -- where tests/golden/floors.npz has one -- next to
the reference's OWN error when its network runs under bf16 autocast (the noise floor of bf16 operands)."""
import os

import numpy as np

_FLOORS = None
REPORT = os.environ.get("PDDM_PARITY_REPORT")  # optional file that collects the lines


def floor(key):
    global _FLOORS
    if _FLOORS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "floors.npz")
        _FLOORS = dict(np.load(path)) if os.path.exists(path) else {}
    v = _FLOORS.get(key)
    return None if v is None else float(v)


def within(name, measured, bound, floor_key=None):
    """assert measured < bound, and say so.  Bounds are 1.5 x the value measured on B200 when the test was written."""
    f = floor(floor_key) if floor_key else None
    line = f"[parity] {name}: measured {measured:.3e}  bound {bound:.1e}" + \
        (f"  reference-under-bf16-autocast {f:.3e}" if f is not None else "")
    print(line, flush=True)
    if REPORT:
        with open(REPORT, "a") as fh:
            fh.write(line + "\n")
    assert measured < bound, line
