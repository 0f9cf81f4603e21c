"""The experimental tile taper (HNB_TILE_TAPER) on the device."""
import numpy as np
import pytest

from tests.helpers import Instance, RefWorld
from tests.test_gpu_update_c5 import _fill, _run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,alive,taper", [(400_000, 380_000, "100"), (400_000, 380_000, "25:1"), (2_200_000, 2_100_000, "100"), (2_200_000, 2_100_000, "50:1")])
def test_c5_with_a_tile_taper(ctx, orc, monkeypatch, rows, alive, taper):
    """Slabs large enough for 2 (>= 114 Ki rows) and 4 (>= 1.8 Mi rows) sub-tiles per tile, so that the taper is active
    (plan_batch only tapers when a tile has more than one sub-tile): every buffer bit-exact after every frame while
    particles die, like tests/test_gpu_update_c5.py. The environment variable is read when the effect source is generated
    (GpuWorld compiles the effect) and when the launch is planned."""
    monkeypatch.setenv("HNB_TILE_TAPER", taper)
    rng = np.random.default_rng(rows % 1000 + len(taper))
    ref = RefWorld(rows, 8, [Instance(0, rows, alive=alive, seed=42)])
    _fill(ref, rng, 0.03, 0.3)
    _run(ctx, orc, ref, 8)
    assert ref.metadata[0].alive_count < alive


def test_many_instances_with_a_tile_taper(ctx, orc, monkeypatch):
    monkeypatch.setenv("HNB_TILE_TAPER", "100")
    rng = np.random.default_rng(3)
    caps = [150_000, 1, 1024, 90_000, 64, 70_000, 7, 100_000]
    alive = [150_000, 1, 1024, 84_321, 0, 52_049, 3, 99_999]
    insts, off = [], 0
    for c, a in zip(caps, alive):
        insts.append(Instance(off, c, alive=a, seed=1000 + off))
        off += c
    ref = RefWorld(off, 8, insts)
    _fill(ref, rng, 0.03, 0.4)
    _run(ctx, orc, ref, 10)
