"""The shared-memory exchanges of every registered multi-pass plan are bank-conflict-free with the pad quantum its
kernels use (tools/bank_sim.py replays every warp-wide write and gather of every exchange), and the plan table of the
simulator matches csrc/wfb_registry.h -- so a plan edited in one place cannot silently go conflicted."""
import re
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import bank_sim  # noqa: E402

PLANS = bank_sim.registered_plans()


@pytest.mark.parametrize("name", sorted(PLANS))
def test_exchanges_are_conflict_free(name):
    n, t, codes, padq, words = PLANS[name]
    worst_write, worst_read = bank_sim.simulate(n, t, codes, padq, words)
    assert worst_write == 1.0 and worst_read == 1.0, (name, worst_write, worst_read)


def test_simulator_table_matches_registry():
    text = (ROOT / "wat-fft_b200" / "csrc" / "wfb_registry.h").read_text()
    decl = {m.group(1): m.group(2) for m in re.finditer(r"using (\w+) = Plan<([^>]+)>;", text)}
    for name, (n, t, codes, _padq, _words) in PLANS.items():
        assert name in decl, name
        args = [int(a, 0) for a in decl[name].replace(" ", "").split(",")]
        assert args[0] == n and args[1] == t and args[2:] == codes, (name, args)
