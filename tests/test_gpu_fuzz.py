"""Seeded random scenarios (tests/fuzz_cases.py): corridors with lane changing (2-4 lanes, off-ramps, forced changes, signals, bounded
acceleration, relaxation, exponential headways) and signalised grids (1-2 lanes, demand and / or pre-seeded populations, shuffled
list order, two vehicle types, leader loops, merge points at every junction exit).  Every step of every case is compared with the
oracle, fp64 bit for bit.  The wider campaigns of tools/fuzz_grid.py / fuzz_lane_change.py (120 + 60 seeds, 200 steps) are clean as well
since the merge points are in: the equal-position "twins" of round 1 came from their absence."""
import pytest

from tests import fuzz_cases
from tests.parity import run_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(1, 25))
def test_random_corridor_with_lane_changing(seed):
    kw = fuzz_cases.corridor_case(seed, n_steps=200)
    st = run_parity(fuzz_cases.synth.corridor(**kw), 200)
    assert st["worst"] == 0.0 and st["overflow"] == 0, (kw, st)


@pytest.mark.parametrize("seed", range(1, 41))
def test_random_grid(seed):
    kw = fuzz_cases.grid_case(seed, n_steps=160)
    st = run_parity(fuzz_cases.synth.grid(**kw), 160)
    assert st["worst"] == 0.0 and st["overflow"] == 0 and st["loops_gpu"] == st["loops_oracle"], (kw, st)
