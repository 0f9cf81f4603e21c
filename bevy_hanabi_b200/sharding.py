"""Index-range sharding of one logical effect instance over the GPUs of a box.

No particle reads another particle in init/update, so GPU g simply owns rows [g*P/G, (g+1)*P/G) of every
column of the slab together with its own alive/dead lists and counters; there is no collective on the data
path. The only cross-shard decisions are made on the host from a handful of integers:
  * how a CPU spawn count for the logical instance is split over the shards (`split_spawn`);
  * the logical instance's totals (`merge_counts`), for reporting.
"""
from __future__ import annotations

from typing import Sequence


def shard_range(total_rows: int, rank: int, world: int) -> tuple[int, int]:
    """[first, end) rows owned by `rank`; sizes differ by at most one and the shards tile [0, total_rows)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    first = total_rows * rank // world
    end = total_rows * (rank + 1) // world
    return first, end


def split_spawn(spawn_count: int, free_slots: Sequence[int]) -> list[int]:
    """Split a logical instance's spawn request over its shards proportionally to their free slots
    (`max_spawn_g`), never exceeding a shard's free slots; the remainder goes to the lowest ranks that still
    have room. The total equals min(spawn_count, sum(free_slots)) — what a single-GPU instance would spawn
    (vfx_init.wgsl:115-137 drops the excess)."""
    total_free = sum(free_slots)
    want = min(max(spawn_count, 0), total_free)
    if want == 0:
        return [0] * len(free_slots)
    out = [min(f, want * f // total_free) for f in free_slots]
    rest = want - sum(out)
    for g in range(len(out)):
        if rest == 0:
            break
        add = min(rest, free_slots[g] - out[g])
        out[g] += add
        rest -= add
    assert rest == 0 and all(o <= f for o, f in zip(out, free_slots))
    return out


def merge_counts(per_shard: Sequence[dict]) -> dict:
    """Logical-instance totals from per-shard metadata readbacks ({'alive_count','max_spawn','capacity',...})."""
    keys = ("capacity", "alive_count", "max_spawn", "instance_count", "particle_counter")
    return {k: sum(int(s.get(k, 0)) for s in per_shard) for k in keys}


def dist_exchange():
    """`exchange(value) -> [value of every shard]` over torch.distributed (one process per GPU): the ONLY cross-shard
    traffic of a sharded instance, one integer per shard and step, and host-side. Uses a CPU tensor on gloo and a device
    tensor on NCCL; without a process group the instance has one shard."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return lambda v: [int(v)]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")

    def exchange(v: int) -> list[int]:
        mine = torch.tensor([int(v)], dtype=torch.int64, device=dev)
        out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(out, mine)
        return [int(t.item()) for t in out]
    return exchange


class ShardedInstance:
    """Shard `rank` of ONE logical effect instance of `total_rows` slots split by index range over `world` contexts (one per
    GPU). The shard is a complete instance of its own — rows [first, end) of the logical slab under
    shard-local indices, its own alive / dead lists and counters (`vfx_init.wgsl:115-137` caps apply per shard: a slot freed
    on shard g is reused on shard g) — so every frame is an ordinary `hnb_simulate`; what makes the shards ONE instance is
    decided on the host from one integer per shard:
      * `step(spawn_count)` reads back this shard's `max_spawn`, exchanges it, and takes its part of the logical spawn
        request (`split_spawn`: proportional to the free slots, remainder to the lowest ranks);
      * `fill_c5` gives the shard the rows of the 1-GPU state it owns (one seed, hashed with the LOGICAL row);
      * `checksum()` hashes the rows under their logical index, so the sum over the shards does not depend on `world`;
      * `counts()` are this shard's part of the logical totals (`merge_counts`).
    The effect must publish counts through the mailbox, i.e. not be compiled with HNB_EFFECT_RELAXED_ORDER.
    """

    def __init__(self, ctx, lowered_effect, total_rows: int, rank: int, world: int, exchange=None, effect=None):
        from . import _native as N, runtime as R
        self._N, self._R = N, R
        self.ctx, self.rank, self.world = ctx, rank, world
        self.total_rows = total_rows
        self.first, self.end = shard_range(total_rows, rank, world)
        self.rows = self.end - self.first
        self.stride = lowered_effect.particle_stride
        self.exchange = exchange or dist_exchange()
        self.slab = ctx.slab_create(self.rows, self.stride)
        self.effect = effect if effect is not None else ctx.effect_compile(lowered_effect)
        ctx.metadata_insert(0, R.initial_metadata(self.rows, 0, self.stride // 4))
        ctx.draw_args_insert(0)
        self._batches = (N.BatchInfo * 1)(N.BatchInfo(0, 0, 0, 0, 0, 1))
        self._prefix = (N.u32 * 1)(0)
        self.last_split: list[int] = [0] * world
        # the ONE integer a step needs from the device — this shard's free slots = rows - instance_count of the previous frame —
        # comes through the count mailbox (pinned host memory the update pass posts into): no copy, no stream synchronisation
        ctx.set_count_mailbox(rows=1, ring=4)
        self._last_epoch = None
        self._alive0 = 0

    def fill_c5(self, seed: int, lifetime_lo: float, lifetime_hi: float) -> None:
        """All slots alive, holding rows [first, end) of the logical instance's counter-based C5 state."""
        R = self._R
        self.ctx.slab_fill_c5(self.slab, 0, self.rows, seed, lifetime_lo, lifetime_hi, logical_first=self.first)
        md = R.initial_metadata(self.rows, 0, self.stride // 4)
        md.alive_count, md.max_spawn = self.rows, 0
        self.ctx.metadata_insert(0, md)
        self._alive0, self._last_epoch = self.rows, None

    def step(self, spawn_count: int, dt: float, time: float, seed: int) -> int:
        """One frame of the logical instance on this shard; returns the number of spawns this shard was given."""
        N, R, ctx = self._N, self._R, self.ctx
        mine = 0
        if spawn_count > 0:
            frees = self.exchange(self.free_slots())
            self.last_split = split_spawn(spawn_count, frees)
            mine = self.last_split[self.rank]
        ctx.upload_spawners([R.make_spawner(spawn=mine, seed=seed)])
        ctx.upload_batches_raw(self._batches, 1, self._prefix, 1)
        ctx.set_sim_params(dt, time, 1)
        ctx.simulate([N.BatchLaunch.make(self.effect, self.slab, 0, mine)])
        self._last_epoch = ctx.last_epoch()
        return mine

    def free_slots(self) -> int:
        """Slots this shard can still spawn into = what `max_spawn` will be at the next indirect pass (vfx_indirect.wgsl:66-70:
        capacity - alive_count), from the previous frame's mailbox word."""
        alive = self._alive0 if self._last_epoch is None else self.ctx.mailbox_count(self._last_epoch, 0)
        return self.rows - alive

    def counts(self) -> dict:
        m = self.ctx.read_metadata(0)
        return {"capacity": m.capacity, "alive_count": m.alive_count, "max_spawn": m.max_spawn,
                "instance_count": self.ctx.read_draw_args(0).instance_count, "particle_counter": m.particle_counter}

    def checksum(self) -> int:
        return self.ctx.slab_checksum(self.slab, 0, self.rows, index_base=self.first)

    def close(self) -> None:
        self.ctx.slab_destroy(self.slab)
