import sys, time
sys.path.insert(0, "/root/repo")
import torch
import bevy_hanabi_b200 as hb
from bevy_hanabi_b200 import _native as N, recipes, runtime as R
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = hb.Context(0, stream.cuda_stream)
P = 64<<20
slab = ctx.slab_create(P, 32); ctx.slab_fill_c5(slab, 0, P, 42, 1e9, 1e9)
md = R.initial_metadata(P, 0, 8); md.alive_count = P; md.max_spawn = 0
ctx.metadata_insert(0, md); ctx.draw_args_insert(0)
fx = ctx.effect_compile(recipes.c5_lowered())
sp = (N.Spawner*1)(R.make_spawner(seed=42)); bi = (N.BatchInfo*1)(N.BatchInfo(0,0,0,0,0,1)); pre=(N.u32*1)(0)
ctx.upload_spawners_raw(sp,1); ctx.upload_batches_raw(bi,1,pre,1); ctx.set_sim_params(1/60,0,1)
la = (N.BatchLaunch*1)(N.BatchLaunch.make(fx, slab, 0, 0))
print("idle clock:", [round(ctx.measure_sm_mhz(50)) for _ in range(3)])
for _ in range(30): ctx.simulate_raw(la,1)
print("clock right after 30 queued steps:", round(ctx.measure_sm_mhz(50)))
for gap in (0.0, 0.0005, 0.002, 0.01, 0.1):
    for _ in range(30): ctx.simulate_raw(la,1)
    ctx.sync(); time.sleep(gap)
    print(f"after sync + {gap*1e3:.1f} ms idle: clock {round(ctx.measure_sm_mhz(20))} MHz; then", end=" ")
    ctx.simulate_raw(la,1); print(round(ctx.measure_sm_mhz(20)), end=" ")
    ctx.simulate_raw(la,1); print(round(ctx.measure_sm_mhz(20)), end=" ")
    ctx.simulate_raw(la,1); print(round(ctx.measure_sm_mhz(20)))
