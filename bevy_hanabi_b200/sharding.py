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
