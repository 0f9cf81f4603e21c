"""BASELINE.json's C2 and C3 at their quoted sizes (first green on a B200 in round 2, profiles/r2_pending_tests_first_gpu_run.txt).

tests/test_gpu_effects.py runs the same effects at 4096 / 8192 particles; tests/test_gpu_fullsize.py covers C4 and C5 at
full size through checksums. These two compare EVERY record with the numpy interpreter at the sizes BASELINE.json names
(the interpreter needs a few seconds per frame at 1 Mi particles)."""
import numpy as np
import pytest

from bevy_hanabi_b200 import graph as G
from oracle.hanabi_oracle import pcg_hash
from tests.helpers import Instance, RefWorld
from tests.test_gpu_effects import _firework_trails, _force_field, _run

pytestmark = pytest.mark.gpu


def test_c2_firework_at_32768(ctx, orc):
    """configs[1]: "firework.rs effect, 32768 capacity": bursts into recycled slots, IEEE-exact, zero tolerance."""
    asset = _firework_trails(32768)
    _, size, _ = asset.particle_layout()
    ref = RefWorld(32768, size // 4, [Instance(0, 32768, alive=0)], dt=1.0 / 20.0)
    seeds = lambda f: [int(pcg_hash(np.array([0x4321 + f], dtype=np.uint32))[0])]
    _run(ctx, orc, asset, ref, 50, lambda f: [9000 if f % 20 == 0 else 150], seeds=seeds, check_every=5)
    assert ref.metadata[0].particle_counter > 30000 and 0 < ref.metadata[0].alive_count < 32768


def test_c3_force_field_at_1m(ctx, orc):
    """configs[2]: "force_field.rs: 1M particles": 1e-5 of the attribute's magnitude per step (sphere sampling and the
    ConformToSphere modifiers go through libm), every integer structure exact."""
    n = 1 << 20
    asset = _force_field(n)
    _, size, _ = asset.particle_layout()
    ref = RefWorld(n, size // 4, [Instance(0, n, alive=0, seed=77)])
    props = [{"attraction_accel": 18.0, "repulsor_position": G.Vec3(0.25, 0.5, 0.1)}]
    _run(ctx, orc, asset, ref, 6, lambda f: [n - 4096 if f == 0 else 500], rtol=1e-5, props=props)
    assert ref.metadata[0].alive_count > (n >> 1)
