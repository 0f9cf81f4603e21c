"""Performance matrix beyond the headline bench line: the same kernels on the other BASELINE configurations.

Every row reports the hnb_update launch time (CUDA events inside the context, as in bench.py), the algorithmic
traffic (8 + 2*stride bytes per UPDATED particle) and the fraction of the measured HBM peak; init
rows report whole frames (init + bookkeeping + update of the newly spawned particles) because the init kernel is
not timed separately.  Usage: python tools/perf_matrix.py [scenario ...]
"""
import json
import os
import sys

sys.path.insert(0, "/root/repo")
import numpy as np
import torch

import bevy_hanabi_b200 as hb
from bevy_hanabi_b200 import _native as N, graph as G, recipes, runtime as R

A = G.Attribute
PEAK = 6581.9
try:
    PEAK = float(json.load(open("/root/repo/MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def report(name, ms, bytes_, extra=""):
    gbs = bytes_ / (ms * 1e-3) / 1e9
    print(f"{name:44s} {ms:8.4f} ms  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f} % of {PEAK:.0f}  {extra}", flush=True)


def timed_update(ctx, launches, steps):
    ctx.sync()
    ctx.enable_kernel_timing(True)
    ctx.kernel_time_ms()
    for _ in range(steps):
        ctx.simulate(launches)
    ms, k = ctx.kernel_time_ms()
    ctx.enable_kernel_timing(False)
    return ms / k


def frame_ms(ctx, launches, steps=1):
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ctx.simulate(launches)
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def single_instance(ctx, capacity, stride, alive=0, spawn=0, seed=42):
    md = R.initial_metadata(capacity, 0, stride // 4)
    md.alive_count, md.max_spawn = alive, capacity - alive
    ctx.metadata_insert(0, md)
    ctx.draw_args_insert(0)
    ctx.upload_spawners([R.make_spawner(spawn=spawn, seed=seed)])
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, 1)], [0])
    ctx.set_sim_params(1 / 60, 0.0, 1)


def c5_update():
    P = 64 << 20
    ctx = hb.Context(0, stream.cuda_stream)
    slab = ctx.slab_create(P, 32)
    ctx.slab_fill_c5(slab, 0, P, 42, 1e9, 1e9)
    single_instance(ctx, P, 32, alive=P)
    la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered()), slab, 0, 0)]
    for _ in range(5):
        ctx.simulate(la)
    ms = timed_update(ctx, la, 30)
    report("C5 64Mi update, nobody dies", ms, 72 * P)
    ctx.close()


def c5_dying():
    """Lifetimes U(0, 0.5 s): ~3 % of the survivors die every step; dead-stack pushes + compaction at work."""
    P = 64 << 20
    ctx = hb.Context(0, stream.cuda_stream)
    slab = ctx.slab_create(P, 32)
    ctx.slab_fill_c5(slab, 0, P, 42, 0.0, 0.5)
    single_instance(ctx, P, 32, alive=P)
    la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered()), slab, 0, 0)]
    alive = P
    for step in range(12):
        ms = timed_update(ctx, la, 1)
        after = ctx.read_metadata(0).alive_count
        if step in (0, 1, 5, 11):
            report(f"C5 64Mi dying, step {step}: {alive >> 10} Ki -> {after >> 10} Ki", ms, 72 * alive, f"{100 * (alive - after) / max(alive, 1):.1f} % died")
        alive = after
    ctx.close()


def c4_topology():
    n_inst, cap = 1024, 65536
    P = n_inst * cap
    ctx = hb.Context(0, stream.cuda_stream)
    slab = ctx.slab_create(P, 32)
    sp = []
    for i in range(n_inst):
        ctx.slab_fill_c5(slab, i * cap, cap, 7000 + i, 1e9, 1e9)
        m = R.initial_metadata(cap, i, 8)
        m.alive_count, m.max_spawn = cap, 0
        ctx.metadata_insert(i, m)
        ctx.draw_args_insert(i)
        sp.append(R.make_spawner(seed=i, effect_metadata_index=i, draw_indirect_index=i, slab_offset=i * cap))
    ctx.upload_spawners(sp)
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, n_inst)], [0] * n_inst)
    ctx.set_sim_params(1 / 60, 0.0, n_inst)
    la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered()), slab, 0, 0)]
    for _ in range(5):
        ctx.simulate(la)
    ms = timed_update(ctx, la, 30)
    report("C4 shape: 1024 instances x 65536, one batch", ms, 72 * P, f"frame {frame_ms(ctx, la, 20):.4f} ms")
    # the same with one spawn per instance per frame (init + update): needs free slots -> kill a few first
    ctx.close()


def _burst(name, asset, P, spawn, props=None, steps=10, fast_math=False, sector_planes=False):
    ctx = hb.Context(0, stream.cuda_stream)
    fx = asset.generate(fast_math=fast_math, sector_planes=sector_planes)
    stride = fx.particle_stride
    slab = ctx.slab_create(P, stride, sector_planes=sector_planes)
    effect = ctx.effect_compile(fx)
    if props is not None:
        ctx.upload_properties(effect, 0, props)
    md = R.initial_metadata(P, 0, stride // 4, properties_array_index=0 if props is not None else N.INVALID)
    ctx.metadata_insert(0, md)
    ctx.draw_args_insert(0)
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, 1)], [0])
    ctx.set_sim_params(1 / 60, 0.0, 1)
    # warm the kernels on a throw-away frame pair, then reset the instance
    ctx.upload_spawners([R.make_spawner(spawn=1024, seed=1)])
    ctx.simulate([N.BatchLaunch.make(effect, slab, 0, 1024)])
    ctx.sync()
    ctx.slab_reset_rows(slab, 0, P)
    ctx.metadata_insert(0, md)
    ctx.upload_spawners([R.make_spawner(spawn=spawn, seed=7)])
    t_burst = frame_ms(ctx, [N.BatchLaunch.make(effect, slab, 0, spawn)])
    alive = ctx.read_metadata(0).alive_count
    ctx.upload_spawners([R.make_spawner(spawn=0, seed=8)])
    la = [N.BatchLaunch.make(effect, slab, 0, 0)]
    for _ in range(3):
        ctx.simulate(la)
    ms = timed_update(ctx, la, steps)
    alive2 = ctx.read_metadata(0).alive_count
    upd_bytes = (8 + 2 * stride) * alive
    report(f"{name}: update, {alive2 >> 10} Ki alive, stride {stride}", ms, (8 + 2 * stride) * alive2)
    t_plain = frame_ms(ctx, la, 5)
    init_bytes = (stride + 8) * alive
    report(f"{name}: burst frame (init {alive >> 10} Ki + update)", t_burst, init_bytes + upd_bytes, f"init alone ~{t_burst - t_plain:.4f} ms = {init_bytes / max(t_burst - t_plain, 1e-6) / 1e6:.0f} GB/s")
    ctx.close()


def c5_init_burst():
    w = G.ExprWriter()
    asset = (G.EffectAsset(64 << 20, w.module, name="c5_spawned")
             .init(G.SetAttributeModifier(A.POSITION, w.rand(G.VEC3) * w.lit(2.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.VELOCITY, w.rand(G.VEC3) * w.lit(2.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.AGE, w.lit(0.)))
             .init(G.SetAttributeModifier(A.LIFETIME, w.lit(1e9)))
             .update(G.AccelModifier(w.lit(G.Vec3(0., -9.8, 0.))))
             .update(G.LinearDragModifier(w.lit(0.5))))
    _burst("C5 recipe, 32Mi burst", asset, 64 << 20, 32 << 20)


def c2_trails():
    from tests.test_gpu_effects import _firework_trails
    _burst("C2 trails, 32Mi burst", _firework_trails(40 << 20), 40 << 20, 32 << 20)


def c3_force_field():
    from tests.test_gpu_effects import _force_field
    for P in (1 << 20, 16 << 20):
        asset = _force_field(P)
        _burst(f"C3 force field, {P >> 20}Mi burst", asset, P, P, props=asset.serialize_properties())


def c3_fast_math():
    from tests.test_gpu_effects import _force_field
    asset = _force_field(16 << 20)
    _burst("C3 force field FAST_MATH, 16Mi burst", asset, 16 << 20, 16 << 20, props=asset.serialize_properties(), fast_math=True)


def many_batches():
    """Typical game frame: many effect assets -> many batches -> one init/update launch each."""
    for nb, cap in ((64, 16 << 10), (64, 256 << 10), (256, 4 << 10)):
        ctx = hb.Context(0, stream.cuda_stream)
        effect = ctx.effect_compile(recipes.c5_lowered())
        sp, bis, la = [], [], []
        for b in range(nb):
            slab = ctx.slab_create(cap, 32)
            ctx.slab_fill_c5(slab, 0, cap, 100 + b, 1e9, 1e9)
            m = R.initial_metadata(cap, b, 8)
            m.alive_count, m.max_spawn = cap, 0
            ctx.metadata_insert(b, m)
            ctx.draw_args_insert(b)
            sp.append(R.make_spawner(seed=b, effect_metadata_index=b, draw_indirect_index=b, slab_offset=0))
            bis.append(N.BatchInfo(0, 0, b, 0, b, 1))
            la.append(N.BatchLaunch.make(effect, slab, b, 0))
        ctx.upload_spawners(sp)
        ctx.upload_batches(bis, [0] * nb)
        ctx.set_sim_params(1 / 60, 0.0, nb)
        for _ in range(5):
            ctx.simulate(la)
        ms = frame_ms(ctx, la, 30)
        report(f"{nb} batches x {cap >> 10} Ki particles: frame", ms, 72 * nb * cap, f"{ms * 1e3 / nb:.1f} us per batch")
        ctx.close()


def churn(sector_planes: bool = False, slot_order: bool = False):
    """Steady-state churn: constant spawn rate into recycled slots until the alive list is a random-looking permutation
    of the slab (survivors keep their relative order, new particles land in whatever slots died). The gathers then touch
    scattered 16-byte plane elements (half-used 32-byte sectors): the honest number for long-running effects, unlike the
    freshly filled slabs of the other rows."""
    from tests.test_gpu_scene import _drifting_sparks
    P = 16 << 20
    ctx = hb.Context(0, stream.cuda_stream)
    asset = _drifting_sparks(P)          # lifetimes U(0.2, 0.9) s
    fx = asset.generate(sector_planes=sector_planes, slot_order=slot_order)
    stride = fx.particle_stride
    slab = ctx.slab_create(P, stride, sector_planes=sector_planes)
    effect = ctx.effect_compile(fx)
    ctx.metadata_insert(0, R.initial_metadata(P, 0, stride // 4))
    ctx.draw_args_insert(0)
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, 1)], [0])
    dt, rate = 1 / 60, P // 40           # ~33 frames mean life -> the population settles around 0.8 P
    for f in range(240):                 # 4 s of simulated time: every slot has been recycled several times
        ctx.set_sim_params(dt, f * dt, 1)
        ctx.upload_spawners([R.make_spawner(spawn=rate, seed=1000 + f)])
        ctx.simulate([N.BatchLaunch.make(effect, slab, 0, rate)])
    alive = ctx.read_metadata(0).alive_count
    ind = ctx.slab_download_indirect(slab, 0, 1 << 16)
    col = ctx.read_metadata(0).indirect_write_index
    jumps = np.abs(np.diff(ind[:, col].astype(np.int64)))
    la = [N.BatchLaunch.make(effect, slab, 0, rate)]
    ctx.sync(); ctx.enable_kernel_timing(True); ctx.kernel_time_ms()
    for f in range(240, 250):
        ctx.set_sim_params(dt, f * dt, 1)
        ctx.upload_spawners([R.make_spawner(spawn=rate, seed=1000 + f)])
        ctx.simulate(la)
    ms, k = ctx.kernel_time_ms()
    ctx.enable_kernel_timing(False)
    alive2 = ctx.read_metadata(0).alive_count
    report(f"churn steady state{' (sector planes)' if sector_planes else ''}{' (SLOT ORDER)' if slot_order else ''}: {alive2 >> 10} Ki of {P >> 10} Ki alive, stride {stride}", ms / k, (8 + 2 * stride) * alive2,
           f"median |slot jump| between consecutive alive-list entries {np.median(jumps):.0f} (1 = identity order)")
    ctx.close()


def churn_slot():
    """The same with HNB_EFFECT_SLOT_ORDER: the update walks the slots, not the (permuted) alive list."""
    churn(slot_order=True)


def c5_slot():
    """C5 in slot order, nobody dies: the headline workload without the alive-list read (68 B instead of 72 B per particle;
    reported against the same 72 B so that the rows compare)."""
    for mi in (8, 64):
        P = mi << 20
        ctx = hb.Context(0, stream.cuda_stream)
        slab = ctx.slab_create(P, 32)
        ctx.slab_fill_c5(slab, 0, P, 42, 1e9, 1e9)
        single_instance(ctx, P, 32, alive=P)
        la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered(slot_order=True)), slab, 0, 0)]
        for _ in range(10):
            ctx.simulate(la)
        fr = min(frame_ms(ctx, la, 100) for _ in range(3))
        k = timed_update(ctx, la, 30)
        assert ctx.read_metadata(0).alive_count == P
        report(f"C5 {mi:2d}Mi SLOT ORDER frame chain", fr, 72 * P, f"isolated update kernel {k:.4f} ms = {68 * P / k / 1e6:.0f} GB/s of its own 68 B per particle")
        ctx.close()
    # dying: the population thins out but every access stays inside contiguous spans
    P = 64 << 20
    ctx = hb.Context(0, stream.cuda_stream)
    slab = ctx.slab_create(P, 32)
    ctx.slab_fill_c5(slab, 0, P, 42, 0.0, 0.5)
    single_instance(ctx, P, 32, alive=P)
    la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered(slot_order=True)), slab, 0, 0)]
    alive = P
    for step in range(12):
        ms = timed_update(ctx, la, 1)
        after = ctx.read_metadata(0).alive_count
        if step in (0, 1, 5, 11):
            report(f"C5 64Mi SLOT ORDER dying, step {step}: {alive >> 10} Ki -> {after >> 10} Ki", ms, 72 * alive, f"{100 * (alive - after) / max(alive, 1):.1f} % died")
        alive = after
    ctx.close()


def churn_sector():
    """The same with HNB_SLAB_SECTOR_PLANES (32-byte-wide columns): one full sector per gathered record."""
    churn(sector_planes=True)


def fresh_sector():
    """Cost of sector planes when access IS coalesced: C5 recipe burst + update, like `c5_init`, on a sector slab."""
    w = G.ExprWriter()
    asset = (G.EffectAsset(64 << 20, w.module, name="c5_spawned_sector")
             .init(G.SetAttributeModifier(A.POSITION, w.rand(G.VEC3) * w.lit(2.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.VELOCITY, w.rand(G.VEC3) * w.lit(2.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.AGE, w.lit(0.)))
             .init(G.SetAttributeModifier(A.LIFETIME, w.lit(1e9)))
             .update(G.AccelModifier(w.lit(G.Vec3(0., -9.8, 0.))))
             .update(G.LinearDragModifier(w.lit(0.5))))
    _burst("C5 recipe on SECTOR planes, 32Mi burst", asset, 64 << 20, 32 << 20, sector_planes=True)


def frame_chain():
    """Whole frames (bookkeeping + update, state resident, no table changes) back to back, with and without programmatic
    dependent launch: what one step costs beyond its update kernel, at the shard sizes of the strong-scaling runs."""
    import os
    for mi in (1, 2, 4, 8, 16, 64):
        P = mi << 20
        for pdl in ("0", "1"):
            os.environ["HNB_PDL"] = pdl
            ctx = hb.Context(0, stream.cuda_stream)
            slab = ctx.slab_create(P, 32)
            ctx.slab_fill_c5(slab, 0, P, 42, 1e9, 1e9)
            single_instance(ctx, P, 32, alive=P)
            la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered()), slab, 0, 0)]
            for _ in range(10):
                ctx.simulate(la)
            fr = min(frame_ms(ctx, la, 200) for _ in range(3))
            k = timed_update(ctx, la, 50)
            report(f"C5 {mi:2d}Mi frame chain, HNB_PDL={pdl}", fr, 72 * P, f"isolated update kernel {k:.4f} ms; frame - kernel = {1e3 * (fr - k):+.1f} us")
            ctx.close()
    os.environ.pop("HNB_PDL", None)


def chunks_sweep():
    """Tile size (HNB_TILE_CHUNKS sub-tiles of 128 rows) and CTAs per SM under the pipelined frame chain."""
    import os
    for mi in (2, 4, 8, 16, 32):
        P = mi << 20
        for defines in ("", "HNB_MIN_BLOCKS=4"):
            for chunks in ("1", "2", "4"):
                os.environ["HNB_TILE_CHUNKS"] = chunks
                if defines:
                    os.environ["HNB_DEFINES"] = defines
                else:
                    os.environ.pop("HNB_DEFINES", None)
                ctx = hb.Context(0, stream.cuda_stream)
                slab = ctx.slab_create(P, 32)
                ctx.slab_fill_c5(slab, 0, P, 42, 1e9, 1e9)
                single_instance(ctx, P, 32, alive=P)
                la = [N.BatchLaunch.make(ctx.effect_compile(recipes.c5_lowered()), slab, 0, 0)]
                for _ in range(10):
                    ctx.simulate(la)
                fr = min(frame_ms(ctx, la, 200) for _ in range(3))
                report(f"C5 {mi:2d}Mi frame chain, chunks={chunks} {defines}", fr, 72 * P)
                ctx.close()
    os.environ.pop("HNB_TILE_CHUNKS", None)
    os.environ.pop("HNB_DEFINES", None)


def interop():
    """Device-resident export of the reference layouts (what a renderer binds): SoA planes -> AoS records and
    {ping,pong,dead} columns -> interleaved rows, device to device; bytes = read + written."""
    for stride, P in ((32, 64 << 20), (48, 32 << 20), (20, 32 << 20)):
        ctx = hb.Context(0, stream.cuda_stream)
        slab = ctx.slab_create(P, stride)
        buf = ctx.device_alloc(P * stride)
        ibuf = ctx.device_alloc(P * 12)
        for name, fn, nbytes in (("export AoS", lambda: ctx.slab_export_aos_device(slab, 0, P, buf), 2 * P * stride),
                                 ("import AoS", lambda: ctx.slab_import_aos_device(slab, 0, P, buf), 2 * P * stride),
                                 ("export indirect rows", lambda: ctx.slab_export_indirect_device(slab, 0, P, ibuf), 2 * P * 12),
                                 ("import indirect rows", lambda: ctx.slab_import_indirect_device(slab, 0, P, ibuf), 2 * P * 12)):
            for _ in range(3):
                fn()
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                fn()
            e1.record(stream)
            e1.synchronize()
            report(f"interop {name}, stride {stride}, {P >> 20} Mi rows", e0.elapsed_time(e1) / 10, nbytes)
        ctx.device_free(buf)
        ctx.device_free(ibuf)
        ctx.close()


def c2_small():
    """BASELINE configs[1] at ITS size: firework trails, 32768 slots (48-byte records). The whole population is 1.5 MB: a
    frame is pure fixed cost (launch chain + dependent-load latency), reported in microseconds per frame."""
    from tests.test_gpu_effects import _firework_trails
    P = 32768
    ctx = hb.Context(0, stream.cuda_stream)
    fx = _firework_trails(P).generate()
    slab = ctx.slab_create(P, fx.particle_stride)
    effect = ctx.effect_compile(fx)
    ctx.metadata_insert(0, R.initial_metadata(P, 0, fx.particle_stride // 4))
    ctx.draw_args_insert(0)
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, 1)], [0])
    ctx.set_sim_params(1 / 600, 0.0, 1)     # short steps: nobody expires during the measurement
    ctx.upload_spawners([R.make_spawner(spawn=30000, seed=7)])
    t_burst = frame_ms(ctx, [N.BatchLaunch.make(effect, slab, 0, 30000)])
    ctx.upload_spawners([R.make_spawner(spawn=0, seed=8)])
    la = [N.BatchLaunch.make(effect, slab, 0, 0)]
    for _ in range(10):
        ctx.simulate(la)
    fr = min(frame_ms(ctx, la, 200) for _ in range(3))
    k = timed_update(ctx, la, 50)
    alive = ctx.read_metadata(0).alive_count
    report(f"C2 firework @ 32768 slots: steady frame, {alive} alive", fr, (8 + 2 * 48) * alive, f"{fr * 1e3:.1f} us per frame (update kernel alone {k * 1e3:.1f} us); first burst frame of 30000 spawns (cold) {t_burst * 1e3:.1f} us")
    # a frame that also spawns: 100 particles per frame into free slots
    ctx.upload_spawners([R.make_spawner(spawn=8, seed=9)])
    la8 = [N.BatchLaunch.make(effect, slab, 0, 8)]
    for _ in range(5):
        ctx.simulate(la8)
    fr8 = min(frame_ms(ctx, la8, 100) for _ in range(3))
    report("C2 firework @ 32768 slots: frame with 8 spawns (init + bookkeeping + update)", fr8, (8 + 2 * 48) * alive, f"{fr8 * 1e3:.1f} us per frame")
    ctx.close()


def c4_recipe():
    """BASELINE configs[3] as stated: instancing.rs recipe, 1024 instances x 65536 slots in ONE batch, filled
    through the real init kernel; update-only frames, then frames that spawn one particle per instance (init with a
    depth-10 prefix search over 1024 instances + bookkeeping of 1024 instances + update)."""
    from tests.test_gpu_config_sizes import _instancing
    n_inst, cap = 1024, 65536
    ctx = hb.Context(0, stream.cuda_stream)
    fx = _instancing(cap).generate()
    stride = fx.particle_stride
    slab = ctx.slab_create(n_inst * cap, stride)
    effect = ctx.effect_compile(fx)
    for i in range(n_inst):
        ctx.metadata_insert(i, R.initial_metadata(cap, i, stride // 4))
        ctx.draw_args_insert(i)
    fill = cap - 1024
    mk = lambda spawn: [R.make_spawner(spawn=spawn, seed=1000 + i, effect_metadata_index=i, draw_indirect_index=i, slab_offset=i * cap) for i in range(n_inst)]
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, n_inst)], [i * fill for i in range(n_inst)])
    ctx.set_sim_params(1 / 60, 0.0, n_inst)
    ctx.upload_spawners(mk(fill))
    t_fill = frame_ms(ctx, [N.BatchLaunch.make(effect, slab, 0, n_inst * fill)])
    alive = n_inst * fill
    report(f"C4 recipe: fill frame (init {alive >> 10} Ki over 1024 instances + update)", t_fill, (stride + 8) * alive + (8 + 2 * stride) * alive)
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, n_inst)], [0] * n_inst)
    ctx.upload_spawners(mk(0))
    la = [N.BatchLaunch.make(effect, slab, 0, 0)]
    for _ in range(5):
        ctx.simulate(la)
    fr = min(frame_ms(ctx, la, 30) for _ in range(2))
    k = timed_update(ctx, la, 20)
    report(f"C4 recipe: update-only frame, {alive >> 10} Ki alive in 1024 instances", fr, (8 + 2 * stride) * alive, f"update kernel alone {k:.4f} ms")
    ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, n_inst)], list(range(n_inst)))
    ctx.upload_spawners(mk(1))
    la1 = [N.BatchLaunch.make(effect, slab, 0, n_inst)]
    for _ in range(5):
        ctx.simulate(la1)
    fr1 = min(frame_ms(ctx, la1, 30) for _ in range(2))
    alive1 = sum(ctx.read_metadata(i).alive_count for i in (0, 511, 1023)) // 3 * n_inst
    report(f"C4 recipe: spawn 1 / instance / step frame (init 1024 + update {alive1 >> 10} Ki)", fr1, (8 + 2 * stride) * alive1, f"+{(fr1 - fr) * 1e3:.1f} us over the update-only frame")
    ctx.close()


def c3_chain():
    """C3 at its BASELINE size (1 Mi) as a frame chain (what a running effect costs per frame), strict and fast-math."""
    from tests.test_gpu_effects import _force_field
    P = 1 << 20
    for fast in (False, True):
        asset = _force_field(P)
        ctx = hb.Context(0, stream.cuda_stream)
        fx = asset.generate(fast_math=fast)
        slab = ctx.slab_create(P, fx.particle_stride)
        effect = ctx.effect_compile(fx)
        ctx.upload_properties(effect, 0, asset.serialize_properties())
        ctx.metadata_insert(0, R.initial_metadata(P, 0, fx.particle_stride // 4, properties_array_index=0))
        ctx.draw_args_insert(0)
        ctx.upload_batches([N.BatchInfo(0, 0, 0, 0, 0, 1)], [0])
        ctx.set_sim_params(1 / 6000, 0.0, 1)
        ctx.upload_spawners([R.make_spawner(spawn=P, seed=7)])
        ctx.simulate([N.BatchLaunch.make(effect, slab, 0, P)])
        ctx.upload_spawners([R.make_spawner(spawn=0, seed=8)])
        la = [N.BatchLaunch.make(effect, slab, 0, 0)]
        for _ in range(10):
            ctx.simulate(la)
        fr = min(frame_ms(ctx, la, 200) for _ in range(3))
        k = timed_update(ctx, la, 50)
        alive = ctx.read_metadata(0).alive_count
        report(f"C3 force field 1Mi frame chain{' FAST_MATH' if fast else ''}, {alive >> 10} Ki alive", fr, (8 + 2 * fx.particle_stride) * alive, f"isolated update kernel {k:.4f} ms")
        ctx.close()


SCENARIOS = {"c2_small": c2_small, "c4_recipe": c4_recipe, "c3_chain": c3_chain, "churn_slot": churn_slot, "c5_slot": c5_slot, "chunks": chunks_sweep, "interop": interop, "frame_chain": frame_chain, "churn": churn, "churn_sector": churn_sector, "fresh_sector": fresh_sector, "many": many_batches, "c5": c5_update, "c5_dying": c5_dying, "c4": c4_topology, "c5_init": c5_init_burst, "c2": c2_trails, "c3": c3_force_field, "c3_fast": c3_fast_math}
if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SCENARIOS)):
        try:
            SCENARIOS[name]()
        except Exception as e:  # keep going: one scenario must not hide the others
            print(f"{name}: FAILED {type(e).__name__}: {str(e)[:400]}", flush=True)
