"""Seeded random scenarios (tests/fuzz_cases.py): corridors with lane changing (2-4 lanes, off-ramps, forced changes, signals, bounded
acceleration, relaxation, exponential headways) and signalised grids (1-2 lanes, demand and / or pre-seeded populations, shuffled
list order, two vehicle types, leader loops).  Every step of every case is compared with the oracle, fp64 bit for bit."""
import pytest

from tests import fuzz_cases
from tests.parity import run_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(1, 13))
def test_random_corridor_with_lane_changing(seed):
    kw = fuzz_cases.corridor_case(seed, n_steps=200)
    st = run_parity(fuzz_cases.synth.corridor(**kw), 200)
    assert st["worst"] == 0.0 and st["overflow"] == 0, (kw, st)


@pytest.mark.parametrize("seed", range(1, 13))
def test_random_grid(seed):
    kw = fuzz_cases.grid_case(seed, n_steps=160)
    st = run_parity(fuzz_cases.synth.grid(**kw), 160)
    assert st["worst"] == 0.0 and st["overflow"] == 0 and st["loops_gpu"] == st["loops_oracle"], (kw, st)
