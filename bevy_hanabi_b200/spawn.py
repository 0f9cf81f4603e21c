"""Binding of the CPU producers (SpawnerSettings / EffectSpawner::tick, Batcher::push) — reference src/spawn.rs,
src/render/batch.rs — implemented natively in csrc/graph/spawn_batch.cpp."""
from __future__ import annotations

import ctypes as C

from . import _native as N
from ._native import check, lib

P, u32, f32 = C.POINTER, C.c_uint32, C.c_float


class SpawnerSettingsC(C.Structure):
    _fields_ = [("count_lo", f32), ("count_hi", f32), ("spawn_duration_lo", f32), ("spawn_duration_hi", f32), ("period_lo", f32),
                ("period_hi", f32), ("cycle_count", u32), ("starts_active", u32), ("emit_on_start", u32)]


class SpawnerState(C.Structure):
    _fields_ = [("cycle_time", f32), ("cycle_spawn_duration", f32), ("cycle_period", f32), ("cycle_ratio", f32), ("cycle_spawn_count", f32),
                ("completed_cycle_count", u32), ("active", u32), ("has_completed", u32), ("spawn_count", u32)]


class BatchKey(C.Structure):
    _fields_ = [("asset_id", C.c_uint64), ("slab_id", u32), ("pipeline_id", u32), ("property_key", u32), ("parent_slab_id", u32),
                ("uses_gpu_events", u32), ("is_cpu_spawner", u32)]


SPAWN_SIGNATURES = {
    "hnb_spawner_settings_new": (C.c_int32, [f32, f32, f32, f32, f32, f32, u32, P(SpawnerSettingsC)]),
    "hnb_spawner_settings_once": (C.c_int32, [f32, P(SpawnerSettingsC)]),
    "hnb_spawner_settings_rate": (C.c_int32, [f32, P(SpawnerSettingsC)]),
    "hnb_spawner_settings_burst": (C.c_int32, [f32, f32, P(SpawnerSettingsC)]),
    "hnb_effect_spawner_create": (C.c_void_p, [P(SpawnerSettingsC), C.c_uint64]),
    "hnb_effect_spawner_destroy": (None, [C.c_void_p]),
    "hnb_effect_spawner_tick": (C.c_int32, [C.c_void_p, f32, P(u32)]),
    "hnb_effect_spawner_reset": (None, [C.c_void_p]),
    "hnb_effect_spawner_set_active": (None, [C.c_void_p, u32]),
    "hnb_effect_spawner_state": (C.c_int32, [C.c_void_p, P(SpawnerState)]),
    "hnb_effect_sorter_create": (C.c_void_p, []),
    "hnb_effect_sorter_destroy": (None, [C.c_void_p]),
    "hnb_effect_sorter_insert": (None, [C.c_void_p, C.c_uint64, u32, u32, C.c_uint64]),
    "hnb_effect_sorter_sort": (C.c_int32, [C.c_void_p]),
    "hnb_effect_sorter_len": (u32, [C.c_void_p]),
    "hnb_effect_sorter_get": (C.c_uint64, [C.c_void_p, u32]),
    "hnb_batcher_create": (C.c_void_p, []),
    "hnb_batcher_destroy": (None, [C.c_void_p]),
    "hnb_batcher_clear": (None, [C.c_void_p]),
    "hnb_batcher_push": (C.c_int32, [C.c_void_p, P(BatchKey), u32, u32, u32, P(C.c_int32)]),
    "hnb_batcher_finish": (C.c_int32, [C.c_void_p, P(P(N.BatchInfo)), P(u32), P(P(u32)), P(u32), P(u32), u32]),
}
for _n, (_r, _a) in SPAWN_SIGNATURES.items():
    _f = getattr(lib, _n)
    _f.restype, _f.argtypes = _r, _a


def _cv(x):
    return (float(x), float(x)) if not isinstance(x, (tuple, list)) else (float(x[0]), float(x[1]))


class SpawnerSettings:
    """`count`, `spawn_duration`, `period`: a float (CpuValue::Single) or a (lo, hi) pair (CpuValue::Uniform)."""

    def __init__(self, count, spawn_duration, period, cycle_count: int):
        self.c = SpawnerSettingsC()
        (cl, ch), (dl, dh), (pl, ph) = _cv(count), _cv(spawn_duration), _cv(period)
        check(lib.hnb_spawner_settings_new(cl, ch, dl, dh, pl, ph, cycle_count, C.byref(self.c)))

    @classmethod
    def once(cls, count):
        return cls(count, 0.0, 0.0, 1)

    @classmethod
    def rate(cls, rate):
        return cls(rate, 1.0, 1.0, 0)

    @classmethod
    def burst(cls, count, period):
        return cls(count, 0.0, period, 0)

    def with_starts_active(self, v: bool):
        self.c.starts_active = int(v)
        return self

    def with_emit_on_start(self, v: bool):
        self.c.emit_on_start = int(v)
        return self

    def is_once(self):
        return self.c.cycle_count == 1

    def is_forever(self):
        return self.c.cycle_count == 0


class EffectSpawner:
    def __init__(self, settings: SpawnerSettings, rng_seed: int = 0):
        self._h = lib.hnb_effect_spawner_create(C.byref(settings.c), rng_seed)

    def __del__(self):
        try:
            if self._h:
                lib.hnb_effect_spawner_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def tick(self, dt: float) -> int:
        n = u32()
        check(lib.hnb_effect_spawner_tick(self._h, dt, C.byref(n)))
        return n.value

    def reset(self):
        lib.hnb_effect_spawner_reset(self._h)

    @property
    def state(self) -> SpawnerState:
        s = SpawnerState()
        check(lib.hnb_effect_spawner_state(self._h, C.byref(s)))
        return s

    @property
    def active(self) -> bool:
        return bool(self.state.active)

    @active.setter
    def active(self, v: bool):
        lib.hnb_effect_spawner_set_active(self._h, int(v))

    def has_completed(self) -> bool:
        return bool(self.state.has_completed)


class Batcher:
    def __init__(self):
        self._h = lib.hnb_batcher_create()

    def __del__(self):
        try:
            if self._h:
                lib.hnb_batcher_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def clear(self):
        lib.hnb_batcher_clear(self._h)

    def push(self, key: BatchKey, spawner_base: int, slab_offset: int, instance_spawn_count: int) -> int:
        idx = C.c_int32()
        check(lib.hnb_batcher_push(self._h, C.byref(key), spawner_base, slab_offset, instance_spawn_count, C.byref(idx)))
        return idx.value

    def finish(self):
        """-> (list[BatchInfo], prefix sums, per-batch CPU spawn totals)"""
        infos, prefix = P(N.BatchInfo)(), P(u32)()
        nb, np_ = u32(), u32()
        totals = (u32 * 4096)()
        check(lib.hnb_batcher_finish(self._h, C.byref(infos), C.byref(nb), C.byref(prefix), C.byref(np_), totals, 4096))
        bi = [N.BatchInfo.from_buffer_copy(bytes(infos[i])) for i in range(nb.value)]
        return bi, [prefix[i] for i in range(np_.value)], list(totals[:nb.value])


NO_ENTITY = 0xFFFFFFFFFFFFFFFF


class EffectSorter:
    """EffectSorter (batch.rs:476-637): children before parents, then by slab, then by row offset."""

    def __init__(self):
        self._h = lib.hnb_effect_sorter_create()

    def __del__(self):
        try:
            if self._h:
                lib.hnb_effect_sorter_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def insert(self, entity: int, slab_id: int, base_instance: int, parent: int | None = None):
        lib.hnb_effect_sorter_insert(self._h, entity, slab_id, base_instance, NO_ENTITY if parent is None else parent)

    def sort(self):
        check(lib.hnb_effect_sorter_sort(self._h))

    def entities(self) -> list[int]:
        return [lib.hnb_effect_sorter_get(self._h, i) for i in range(lib.hnb_effect_sorter_len(self._h))]
