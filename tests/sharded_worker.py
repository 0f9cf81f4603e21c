"""Worker of tests/test_gpu_sharded.py: one process per shard (torchrun, gloo for the one integer per shard and step that
crosses shards), each driving `bevy_hanabi_b200.sharding.ShardedInstance` on its GPU (rank -> cuda:rank % device_count, so the
test also runs on a one-GPU box, both shards on cuda:0 in separate processes and contexts).

Part A — deaths + respawn, exact parity PER SHARD: a logical instance of P slots (C5 attributes, random init, lifetimes
    U(0.05, 0.3) s) receives a logical spawn request every frame; every shard takes its `split_spawn` part. Each rank runs
    the oracle (numpy interpreter + C bookkeeping) for ITS shard — same boundaries, same per-shard spawn counts, same seed —
    and compares every buffer bit for bit (SURVEY.md §8e "compare each shard against an oracle run configured with the same
    shard boundaries"; caps per shard = vfx_init.wgsl:115-137). The split itself is checked against the logical instance's
    semantics: sum over shards == min(request, free slots of all shards).
Part B — the 1-GPU state split `world` ways: C5 fill (one seed, logical rows), update-only steps without and with deaths; the
    sum of the shard checksums and alive counts equals the UNSHARDED oracle run.
"""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    import bevy_hanabi_b200 as hb
    from bevy_hanabi_b200 import graph as G, recipes
    from bevy_hanabi_b200.sharding import ShardedInstance, merge_counts, shard_range
    from oracle import c_oracle as O
    from oracle.hanabi_oracle import EffectOracle, pcg_hash
    from tests.helpers import Instance, RefWorld
    orc = O.load()
    A = G.Attribute

    def gather(v):
        out = [None] * world
        dist.all_gather_object(out, v)
        return out

    # ---------------- Part A
    P = 200_003
    w = G.ExprWriter()
    asset = (G.EffectAsset(P, w.module, name="c5_respawn")
             .init(G.SetAttributeModifier(A.POSITION, w.rand(G.VEC3) * w.lit(2.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.VELOCITY, w.rand(G.VEC3) * w.lit(2.) - w.lit(1.)))
             .init(G.SetAttributeModifier(A.AGE, w.lit(0.)))
             .init(G.SetAttributeModifier(A.LIFETIME, w.lit(0.05).uniform(w.lit(0.3))))
             .update(G.AccelModifier(w.lit(G.Vec3(0., -9.8, 0.))))
             .update(G.LinearDragModifier(w.lit(0.5))))
    ctx = hb.Context(dev)
    sh = ShardedInstance(ctx, asset.generate(), P, rank, world)
    assert (sh.first, sh.end) == shard_range(P, rank, world)
    ref = RefWorld(sh.rows, sh.stride // 4, [Instance(0, sh.rows, alive=0)])
    eo = EffectOracle(asset, None)
    dt = 1.0 / 60.0
    spawned_total = 0
    for f in range(36):
        request = 150_000 if f == 0 else (90_000 if f % 6 == 0 else 9_000)   # often more than the free slots
        # one spawner seed per shard and frame: shard-local indices repeat across shards, the seed keeps their streams apart
        seed = int(pcg_hash(np.array([0x9000 + f * 64 + rank], dtype=np.uint32))[0])
        free_before = sum(gather(int(ref.metadata[0].max_spawn)))
        mine = sh.step(request, dt, f * dt, seed)
        parts = gather(mine)
        assert parts == sh.last_split and sum(parts) == min(request, free_before), (f, parts, free_before)
        spawned_total += sum(parts)
        ref.sim.time = np.float32(f * dt)
        ref.sim.virtual_time = ref.sim.real_time = ref.sim.time
        ref.set_spawns([mine], [seed])
        eo.frame(ref, orc)
        if f % 7 == 0 or f == 35:
            ctx.sync()
            np.testing.assert_array_equal(ctx.slab_download_aos(sh.slab, 0, sh.rows, sh.stride), ref.particles, err_msg=f"shard {rank} frame {f}: particles")
            np.testing.assert_array_equal(ctx.slab_download_indirect(sh.slab, 0, sh.rows), ref.indirect, err_msg=f"shard {rank} frame {f}: lists")
            np.testing.assert_array_equal(np.frombuffer(bytes(ctx.read_metadata(0)), dtype=np.uint32), ref.metadata_rows()[0], err_msg=f"shard {rank} frame {f}: metadata")
            assert ctx.read_draw_args(0).instance_count == ref.draw[1]
    total = merge_counts(gather(sh.counts()))
    assert total["capacity"] == P and total["particle_counter"] == spawned_total and 0 < total["alive_count"] < P
    assert total["alive_count"] + total["max_spawn"] == P and total["instance_count"] == total["alive_count"]
    assert spawned_total > 2 * P, "slots were recycled"
    sh.close()
    ctx.close()

    # ---------------- Part B
    P2 = 1 << 20
    u32p = C.POINTER(C.c_uint32)
    for lo, hi, steps in ((1e9, 1e9, 4), (0.02, 0.12, 5)):
        ctx = hb.Context(dev)
        sh = ShardedInstance(ctx, recipes.c5_lowered(), P2, rank, world)
        sh.fill_c5(1234, lo, hi)
        for s in range(steps):
            sh.step(0, dt, 0.0, 42)
        sums = gather(sh.checksum())
        counts = merge_counts(gather(sh.counts()))
        # the unsharded run (oracle, threads of this host)
        particles = np.empty((P2, 8), dtype=np.float32)
        indirect = np.zeros((P2, 3), dtype=np.uint32)
        indirect[:, 2] = np.arange(P2, dtype=np.uint32)
        orc.orc_fill_c5(O.ptr(particles), O.ptr(indirect), 0, P2, 1234, lo, hi)
        sim = O.SimParams(dt, 0, dt, 0, dt, 0, 1)
        md = (O.EffectMetadata * 1)()
        md[0].capacity, md[0].alive_count, md[0].max_spawn = P2, P2, 0
        sp = (O.Spawner * 1)()
        sp[0].seed = 42
        draw, prefix, dispatch = np.zeros(5, dtype=np.uint32), np.zeros(1, dtype=np.uint32), np.zeros(3, dtype=np.uint32)
        bi = (O.BatchInfo * 1)(O.BatchInfo(0, 0, 0, 0, 0, 1))
        k = (C.c_float * 4)(0.0, -9.8, 0.0, 0.5)
        flags = np.zeros(P2, dtype=np.uint8)
        for _ in range(steps):
            orc.orc_indirect(C.byref(sim), md, draw.ctypes.data_as(u32p), sp, prefix.ctypes.data_as(u32p), None, 0)
            orc.orc_prefix_sum(bi, 1, prefix.ctypes.data_as(u32p), dispatch.ctypes.data_as(u32p))
            orc.orc_update_c5_parallel(C.byref(sim), draw.ctypes.data_as(u32p), O.ptr(particles), O.ptr(indirect), sp, md, k, O.ptr(flags), 4)
        assert sum(sums) % 2**64 == orc.orc_checksum(O.ptr(particles.view(np.uint32)), 0, P2, 8), f"lifetimes [{lo},{hi}]: whole-instance checksum"
        assert counts["alive_count"] == md[0].alive_count == counts["instance_count"] and counts["max_spawn"] == md[0].max_spawn
        if hi < 1e8:
            assert 0 < counts["alive_count"] < P2
        sh.close()
        ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"shard {rank}/{world} on cuda:{dev}: ok", flush=True)


if __name__ == "__main__":
    main()
