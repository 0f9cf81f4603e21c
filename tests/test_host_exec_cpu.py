"""The GENERATED effect code, executed on the CPU (tests/host_exec.py), against the numpy interpreter of the
expression graph: checks the values the lowered text computes — both passes, every record, the is_alive flag —
without a GPU. IEEE-exact effects must agree bit for bit; effects going through libm within 1e-5 of the attribute's
magnitude (glibc and numpy round sin/cos/acos/pow differently in the last place, like CUDA's libm does)."""
import numpy as np
import pytest

from bevy_hanabi_b200 import graph as G
from oracle.hanabi_oracle import EffectOracle, pcg_hash
from tests.helpers import Instance, RefWorld, assert_float_attributes_close
from tests.host_exec import HostEffect, replay_frame
from tests.test_gpu_effects import _firework_trails, _float_word_mask, _force_field
from tests.test_gpu_ribbons import _ribbon_asset
from tests.test_gpu_scene import _drifting_sparks, _growing_dust

A = G.Attribute


def _world(asset, capacity, seed=5, dt=1.0 / 30.0):
    _, size, _ = asset.particle_layout()
    return RefWorld(capacity, size // 4, [Instance(0, capacity, alive=0, seed=seed)], dt=dt)


def _run(orc, asset, frames, spawns, exact=True, props=None, capacity=2048):
    fx = asset.generate()
    host = HostEffect(fx)
    ref = _world(asset, capacity)
    blob = None
    if props is not None:
        blob = asset.serialize_properties(props)
        ref.metadata[0].properties_array_index = 0
    eo = EffectOracle(asset, {0: props} if props else None)
    mask, fattrs = _float_word_mask(asset)
    checked_init = checked_update = deaths = 0
    for f in range(frames):
        ref.sim.time = np.float32(f) * ref.sim.delta_time
        ref.set_spawns([spawns(f)], [int(pcg_hash(np.array([77 + f], dtype=np.uint32))[0])])
        ih, io, uh, uo, ah, ao = replay_frame(host, eo, ref, orc, blob)
        if exact:
            np.testing.assert_array_equal(ih, io, err_msg=f"frame {f}: init records")
            np.testing.assert_array_equal(uh, uo, err_msg=f"frame {f}: update records")
        else:
            np.testing.assert_array_equal(ih[:, ~mask], io[:, ~mask])
            np.testing.assert_array_equal(uh[:, ~mask], uo[:, ~mask])
            if len(ih):
                assert_float_attributes_close(ih, io, fattrs, 1e-5, f"frame {f}: init")
            if len(uh):
                assert_float_attributes_close(uh, uo, fattrs, 1e-5, f"frame {f}: update")
        np.testing.assert_array_equal(ah, ao, err_msg=f"frame {f}: is_alive")
        checked_init += len(ih)
        checked_update += len(uh)
        deaths += int((~ao).sum())
    return checked_init, checked_update, deaths


def test_trails_generated_code_equals_interpreter(orc):
    ci, cu, deaths = _run(orc, _firework_trails(2048), 30, lambda f: 700 if f % 10 == 0 else 0)
    assert ci > 1500
    assert cu > 10_000 and deaths > 100


@pytest.mark.parametrize("name", ["sparks", "dust", "ribbons"])
def test_scene_assets_generated_code_equals_interpreter(orc, name):
    asset = {"sparks": _drifting_sparks, "dust": _growing_dust, "ribbons": _ribbon_asset}[name](2048)
    ci, cu, deaths = _run(orc, asset, 20, lambda f: 300 if f % 4 == 0 else 11)
    assert ci > 800 and cu > 5000
    if name != "ribbons":
        assert deaths > 0


def test_force_field_generated_code_within_tolerance(orc):
    props = {"attraction_accel": 18.0, "repulsor_position": G.Vec3(0.25, 0.5, 0.1)}
    ci, cu, _ = _run(orc, _force_field(4096), 6, lambda f: 3000 if f == 0 else 20, exact=False, props=props, capacity=4096)
    assert ci > 3000 and cu > 15_000


def test_operator_table_generated_code(orc):
    """Every IEEE-exact operator family in one effect (same graph as the GPU operator test uses for its exact part)."""
    w = G.ExprWriter()
    x, v = w.attr(A.F32_0), w.attr(A.VELOCITY)
    asset = (G.EffectAsset(512, w.module, name="ops")
             .init(G.SetAttributeModifier(A.POSITION, w.rand(G.VEC3) * w.lit(4.) - w.lit(2.)))
             .init(G.SetAttributeModifier(A.VELOCITY, w.rand(G.VEC3) - w.lit(0.5)))
             .init(G.SetAttributeModifier(A.F32_0, w.rand() * w.lit(3.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.AGE, w.lit(0.)))
             .init(G.SetAttributeModifier(A.LIFETIME, w.lit(0.2).uniform(w.lit(0.6))))
             .update(G.SetAttributeModifier(A.F32_1, x.abs().sqrt() + x.floor() - x.fract() * x.ceil() + x.sign() + x.saturate() + x.round()))
             .update(G.SetAttributeModifier(A.F32_2, x.min(w.lit(0.5)).max(w.lit(-0.25)).clamp(w.lit(-0.1), w.lit(0.4)).mix(x, w.lit(0.25))
                                            + x.step(w.lit(0.3)) + x.smoothstep(w.lit(-1.), w.lit(2.)) + (x % w.lit(0.7))))
             .update(G.SetAttributeModifier(A.F32X3_0, v.cross(w.attr(A.POSITION)) + v.normalize() * v.length() - v.abs().min(w.attr(A.POSITION))
                                            + v.x().vec3(v.dot(v), v.distance(w.attr(A.POSITION)))))
             .update(G.SetAttributeModifier(A.U32_0, (x.abs() * w.lit(1000.)).cast(G.UINT) + w.attr(A.ID) * w.attr(A.PARTICLE_COUNTER)))
             .update(G.SetAttributeModifier(A.F32_3, v.gt(w.attr(A.POSITION)).any().cast(G.FLOAT) + v.le(w.attr(A.POSITION)).all().cast(G.FLOAT)
                                            + w.is_alive().cast(G.FLOAT) + w.time() * w.delta_time())))
    ci, cu, deaths = _run(orc, asset, 12, lambda f: 400 if f == 0 else 15, capacity=512)
    assert ci > 400 and cu > 2000 and deaths > 50
