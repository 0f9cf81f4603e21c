"""GPU spawn events: a parent effect emits events from its update pass
(EmitSpawnEventModifier, src/modifier/mod.rs:654-717; append_spawn_events_N, src/lib.rs:976-993), a child effect
consumes them in the NEXT frame's init pass (vfx_init.wgsl:123-129, :166-171) and inherits attributes from the
parent particle (InheritAttributeModifier, src/modifier/attr.rs:173-186).

Like in the reference, the ORDER in which events land in the buffer depends on atomic scheduling, so the event
buffer — and therefore which child slot receives which event — is compared as a multiset; counts are exact.
"""
import numpy as np
import pytest

from bevy_hanabi_b200 import _native as N
from bevy_hanabi_b200 import graph as G
from bevy_hanabi_b200 import runtime as R
from oracle.hanabi_oracle import EffectOracle, Env, Rng, Writer, apply_modifier, pcg_hash
from tests.helpers import Instance, RefWorld

pytestmark = pytest.mark.gpu
A = G.Attribute
EVENT_CAP = 256  # the reference hard-codes 256 events per child (event.rs:266-267)


def _assets():
    wp = G.ExprWriter()
    parent = (G.EffectAsset(512, wp.module, name="rocket")
              .init(G.SetAttributeModifier(A.POSITION, wp.rand(G.VEC3) * wp.lit(8.) - wp.lit(4.)))
              .init(G.SetAttributeModifier(A.VELOCITY, wp.lit(G.Vec3(0., 3., 0.))))
              .init(G.SetAttributeModifier(A.AGE, wp.lit(0.)))
              .init(G.SetAttributeModifier(A.LIFETIME, wp.lit(0.05).uniform(wp.lit(0.4))))
              .update(G.EmitSpawnEventModifier(G.ON_DIE, wp.lit(G.U32(3)), 0)))
    wc = G.ExprWriter()
    child = (G.EffectAsset(2048, wc.module, name="sparks")
             .init(G.InheritAttributeModifier(A.POSITION))
             .init(G.SetAttributeModifier(A.VELOCITY, wc.parent_attr(A.VELOCITY) * wc.lit(-0.5) + (wc.rand(G.VEC3) - wc.lit(0.5))))
             .init(G.SetAttributeModifier(A.AGE, wc.lit(0.)))
             .init(G.SetAttributeModifier(A.LIFETIME, wc.lit(0.3)))
             .init(G.SetAttributeModifier(A.U32_0, wc.parent_attr(A.ID))))
    return parent, child


def _sorted_rows(a):
    return a[np.lexsort(a.T[::-1])]


def test_parent_emits_child_consumes(ctx, orc):
    parent, child = _assets()
    p_fx = parent.generate(num_event_bindings=1)
    c_fx = child.generate(parent=parent)
    assert p_fx.flags & N.EFFECT_EMIT_GPU_SPAWN_EVENTS
    assert c_fx.flags & N.EFFECT_CONSUME_GPU_SPAWN_EVENTS and c_fx.flags & N.EFFECT_READ_PARENT_PARTICLE
    p_stride, c_stride = p_fx.particle_stride, c_fx.particle_stride
    dt = 1.0 / 30.0

    # ---- oracle worlds (one slab each) and oracles
    pw = RefWorld(512, p_stride // 4, [Instance(0, 512, alive=0, seed=11)], dt=dt)
    cw = RefWorld(2048, c_stride // 4, [Instance(0, 2048, alive=0, seed=22)], dt=dt)
    po, co = EffectOracle(parent), EffectOracle(child)
    events = np.zeros(EVENT_CAP, dtype=np.uint32)
    event_count = 0

    # ---- GPU: one context, two slabs, spawner/metadata rows 0 (parent) and 1 (child), two batches
    p_slab, c_slab = ctx.slab_create(512, p_stride), ctx.slab_create(2048, c_stride)
    p_eff, c_eff = ctx.effect_compile(p_fx), ctx.effect_compile(c_fx)
    evbuf = ctx.event_buffer_create(EVENT_CAP)
    ctx.child_info_insert(0, 0, 0)
    md_p = R.initial_metadata(512, 0, p_stride // 4)
    md_p.base_child_index = 0
    md_c = R.initial_metadata(2048, 1, c_stride // 4)
    md_c.global_child_index = 0
    md_c.local_child_index = 0
    ctx.metadata_insert(0, md_p)
    ctx.metadata_insert(1, md_c)
    ctx.draw_args_insert(0)
    ctx.draw_args_insert(1)
    launches = [N.BatchLaunch.make(p_eff, p_slab, 0, 0, emit_events=[evbuf]),
                N.BatchLaunch.make(c_eff, c_slab, 1, 0, parent_slab=p_slab, consume_events=evbuf)]

    spawn_sched = [40, 0, 25, 0, 0, 60, 0, 0, 10, 0, 0, 0, 30, 0, 0, 0, 0, 0]
    total_child_spawned = 0
    for f, spawn in enumerate(spawn_sched):
        seed_p = int(pcg_hash(np.array([100 + f], dtype=np.uint32))[0])
        seed_c = int(pcg_hash(np.array([900 + f], dtype=np.uint32))[0])
        # ----- GPU frame
        ctx.upload_spawners([R.make_spawner(spawn=spawn, seed=seed_p, effect_metadata_index=0, draw_indirect_index=0, slab_offset=0),
                             R.make_spawner(spawn=0, seed=seed_c, effect_metadata_index=1, draw_indirect_index=1, slab_offset=0,
                                            parent_slab_offset=0)])
        ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, 1), N.BatchInfo(0, 0, 1, 0, 1, 1)], [0, 0])
        ctx.set_sim_params(dt, f * dt, 2)
        launches[0] = N.BatchLaunch.make(p_eff, p_slab, 0, spawn, emit_events=[evbuf])
        # events written by the previous frame's update, as the child's init is about to see them
        gpu_event_count = ctx.read_child_info(0).event_count
        gpu_events = ctx.event_buffer_download(evbuf, 0, EVENT_CAP)
        assert gpu_event_count == event_count
        n_valid = min(event_count, EVENT_CAP)
        assert sorted(gpu_events[:n_valid].tolist()) == sorted(events[:n_valid].tolist())
        # make the oracle consume the events in the order the GPU buffer holds them (reference order is unspecified)
        events[:n_valid] = gpu_events[:n_valid]
        ctx.simulate(launches)

        # ----- oracle frame
        pw.sim.time = cw.sim.time = np.float32(f * dt)
        pw.set_spawns([spawn], [seed_p])
        cw.set_spawns([0], [seed_c])
        po.init_pass(pw)
        # child init from events (vfx_init.wgsl:123-171)
        md = cw.metadata[0]
        n = min(n_valid if event_count >= 0 else 0, md.max_spawn)
        if n:
            k = np.arange(n, dtype=np.int64)
            alive_index = md.alive_count + k
            slots = cw.indirect[alive_index, 2].astype(np.int64)
            pidx = slots.astype(np.uint32)
            parent_idx = events[:n].astype(np.uint32)
            parent_rec = po.unpack(pw.particles[parent_idx.astype(np.int64)].copy())
            rec = np.zeros((n, co.stride_words), dtype=np.uint32)
            P = co.unpack(rec)
            env = Env(n, P, cw.sim, Rng(pcg_hash(pidx ^ np.uint32(seed_c))), {}, pidx, (md.particle_counter + k).astype(np.uint32),
                      np.array(list(cw.spawners[0].transform), dtype=np.float32).reshape(3, 4), parent=parent_rec, parent_particle_index=parent_idx)
            wr = Writer()
            for m in child.init_modifiers:
                apply_modifier(m, child.module, env, wr)
            co.pack(P, rec)
            cw.indirect[alive_index, md.indirect_write_index] = pidx
            cw.particles[pidx.astype(np.int64)] = rec
            md.alive_count += n
            md.particle_counter += n
            total_child_spawned += n
        # indirect + prefix for both worlds, events cleared (vfx_indirect.wgsl:38-46)
        event_count = 0
        for w in (pw, cw):
            w.oracle_indirect(orc)
            w.oracle_prefix_sum(orc)
        po.update_pass(pw)
        co.update_pass(cw)
        # append_spawn_events_0 in serial thread order (lib.rs:976-993)
        for channel, counts in po.last_emitted:
            assert channel == 0
            pidx_rows = pw.indirect[:pw.metadata[0].max_update, 1 - pw.metadata[0].indirect_write_index]
            for row in np.nonzero(counts)[0]:
                c = int(counts[row])
                base = min(event_count, EVENT_CAP)
                event_count += c
                capped = min(c, EVENT_CAP - base)
                events[base:base + capped] = pidx_rows[row]

        # ----- compare
        ctx.sync()
        for world, slab, row, stride in ((pw, p_slab, 0, p_stride), (cw, c_slab, 1, c_stride)):
            m_gpu = np.frombuffer(bytes(ctx.read_metadata(row)), dtype=np.uint32)
            m_ref = world.metadata_rows()[0].copy()
            m_ref[5] = row  # indirect_draw_index differs between the single-instance oracle worlds and the shared GPU tables
            for fld in (7, 8, 9, 10):
                m_ref[fld] = m_gpu[fld]
            np.testing.assert_array_equal(m_gpu, m_ref, err_msg=f"frame {f} metadata row {row}")
            assert ctx.read_draw_args(row).instance_count == world.draw[1]
        # parent: fully deterministic -> exact
        np.testing.assert_array_equal(ctx.slab_download_aos(p_slab, 0, 512, p_stride), pw.particles)
        np.testing.assert_array_equal(ctx.slab_download_indirect(p_slab, 0, 512), pw.indirect)
        # child: exact as well, because the oracle consumed the events in the GPU buffer's order
        np.testing.assert_array_equal(ctx.slab_download_aos(c_slab, 0, 2048, c_stride), cw.particles)
        np.testing.assert_array_equal(ctx.slab_download_indirect(c_slab, 0, 2048), cw.indirect)
    assert total_child_spawned > 100, "the scenario must actually spawn children from events"
    assert event_count >= 0
