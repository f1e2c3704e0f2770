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


@pytest.mark.parametrize("n", sorted(bank_sim.C2C_F32_TILES))
def test_c2c_tile_io_is_conflict_free(n):
    """The reads of a staged tile into registers and the writes of the result tile: with T < 32 threads per transform a warp
    touches several rows at once, and in a DENSE tile those rows start on the same banks (what ncu showed as 41 % replayed
    shared-memory wavefronts at N = 256 / 512 split, and what the grouped, padded row copies remove).  The registry's
    default layouts are conflict-free; the dense counter-examples are asserted too, so the model stays honest."""
    t, x, rg_split, rg_il = bank_sim.C2C_F32_TILES[n]
    assert bank_sim.c2c_tile_io(n, t, x, rg_split, "split") == 1.0, n
    assert bank_sim.c2c_tile_io(n, t, x, rg_il, "il") == 1.0, n
    if t < 32:                                   # a dense tile would conflict on the split planes (32-bit accesses) ...
        assert bank_sim.c2c_tile_io(n, t, x, 0, "split") == 2.0 * (16 // t if t < 16 else 1), n
    if t < 16:                                   # ... and below 16 threads per transform on interleaved rows as well
        assert bank_sim.c2c_tile_io(n, t, x, 0, "il") > 1.0, n


def test_c2c_tile_defaults_match_the_variant_file():
    text = (ROOT / "wat-fft_b200" / "csrc" / "wfb_variants_f32_pipe.cu").read_text()
    assert "VTSG(F32_128, 16, 2, 2, 16, 61)" in text                       # N = 128: groups of 2, both layouts
    assert "VTSG(F32_256, 8, 2, 4, 16, 61, -1, 58)" in text                # N = 256: groups of 4 on the split layout only
    assert "VTSG(P32_512, 4, 2, 2, 16, 61, -1, 58)" in text                # N = 512: groups of 2 on the split layout only


@pytest.mark.parametrize("name", sorted(bank_sim.REAL_TILES_GROUPED))
def test_real_tile_io_is_conflict_free(name):
    """k_real_pipe where two thread groups share a shared-memory phase: rows T apart keep the (M+1)-bin side conflict-free,
    and the dense side needs the padded groups (2-way conflicts without them: ncu, 33-40 % replayed wavefronts)."""
    m, t, x, eb = bank_sim.REAL_TILES_GROUPED[name]
    dense, odd = bank_sim.real_tile_io(m, t, x, eb, grouped=True)
    assert dense == 1.0 and odd == 1.0, (name, dense, odd)
    dense_old, odd_old = bank_sim.real_tile_io(m, t, x, eb, grouped=False)
    assert dense_old == 2.0 and odd_old == 1.0, (name, dense_old, odd_old)
