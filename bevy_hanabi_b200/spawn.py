"""Binding of the CPU producers (SpawnerSettings / EffectSpawner::tick, Batcher::push, the EffectSimulation clock) —
reference src/spawn.rs, src/render/batch.rs, src/time.rs — implemented natively in csrc/graph/spawn_batch.cpp and
csrc/graph/sim_clock.cpp."""
from __future__ import annotations

import ctypes as C
import math

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


class SimClockState(C.Structure):
    _fields_ = [("real_elapsed_ns", C.c_uint64), ("real_delta_ns", C.c_uint64), ("virtual_elapsed_ns", C.c_uint64),
                ("virtual_delta_ns", C.c_uint64), ("sim_elapsed_ns", C.c_uint64), ("sim_delta_ns", C.c_uint64),
                ("virtual_effective_speed", C.c_double), ("sim_effective_speed", C.c_double)]


SPAWN_SIGNATURES = {
    "hnb_sim_clock_create": (C.c_void_p, []),
    "hnb_sim_clock_destroy": (None, [C.c_void_p]),
    "hnb_sim_clock_set_virtual_relative_speed": (C.c_int32, [C.c_void_p, C.c_double]),
    "hnb_sim_clock_set_virtual_paused": (None, [C.c_void_p, u32]),
    "hnb_sim_clock_set_max_delta_ns": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "hnb_sim_clock_set_relative_speed": (C.c_int32, [C.c_void_p, C.c_double]),
    "hnb_sim_clock_pause": (None, [C.c_void_p]),
    "hnb_sim_clock_unpause": (None, [C.c_void_p]),
    "hnb_sim_clock_is_paused": (u32, [C.c_void_p]),
    "hnb_sim_clock_was_paused": (u32, [C.c_void_p]),
    "hnb_sim_clock_relative_speed": (C.c_double, [C.c_void_p]),
    "hnb_sim_clock_effective_speed": (C.c_double, [C.c_void_p]),
    "hnb_sim_clock_advance": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "hnb_sim_clock_sim_params": (C.c_int32, [C.c_void_p, u32, P(N.SimParams)]),
    "hnb_sim_params_default": (None, [P(N.SimParams)]),
    "hnb_sim_clock_state": (C.c_int32, [C.c_void_p, P(SimClockState)]),
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

    # accessors of the reference (spawn.rs:362-615); CpuValue<f32> is a float (Single) or a (lo, hi) pair (Uniform)
    @staticmethod
    def _value(lo, hi):
        return lo if lo == hi else (lo, hi)

    def set_emit_on_start(self, v: bool):
        self.c.emit_on_start = int(v)

    def emits_on_start(self) -> bool:
        return bool(self.c.emit_on_start)

    def set_count(self, count):
        self.c.count_lo, self.c.count_hi = _cv(count)

    def with_count(self, count):
        self.set_count(count)
        return self

    def count(self):
        return self._value(self.c.count_lo, self.c.count_hi)

    def set_spawn_duration(self, spawn_duration):
        self.c.spawn_duration_lo, self.c.spawn_duration_hi = _cv(spawn_duration)

    def with_spawn_duration(self, spawn_duration):
        self.set_spawn_duration(spawn_duration)
        return self

    def spawn_duration(self):
        return self._value(self.c.spawn_duration_lo, self.c.spawn_duration_hi)

    def set_period(self, period):
        """spawn.rs:538-545: only finiteness is asserted here (positivity is SpawnerSettings::new's check)."""
        lo, hi = _cv(period)
        if not (math.isfinite(lo) and math.isfinite(hi)):
            raise N.HanabiError(N.HNB_ERR_INVALID_ARG, f"`period` {period!r} has an infinite bound. If upgrading from a previous version, "
                                                        "use `cycle_count = 1` instead for a single-cycle burst.")
        self.c.period_lo, self.c.period_hi = lo, hi

    def with_period(self, period):
        self.set_period(period)
        return self

    def period(self):
        return self._value(self.c.period_lo, self.c.period_hi)

    def set_cycle_count(self, cycle_count: int):
        self.c.cycle_count = int(cycle_count)

    def with_cycle_count(self, cycle_count: int):
        self.set_cycle_count(cycle_count)
        return self

    def cycle_count(self) -> int:
        return self.c.cycle_count

    def set_starts_active(self, v: bool):
        self.c.starts_active = int(v)

    def starts_active(self) -> bool:
        return bool(self.c.starts_active)

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

    def with_active(self, active: bool) -> "EffectSpawner":   # spawn.rs:723
        lib.hnb_effect_spawner_set_active(self._h, int(active))
        return self

    # per-cycle accessors of the reference (spawn.rs:730-800)
    def cycle_time(self) -> float: return self.state.cycle_time
    def cycle_spawn_duration(self) -> float: return self.state.cycle_spawn_duration
    def cycle_period(self) -> float: return self.state.cycle_period
    def cycle_ratio(self) -> float: return self.state.cycle_ratio
    def cycle_spawn_count(self) -> float: return self.state.cycle_spawn_count
    def completed_cycle_count(self) -> int: return self.state.completed_cycle_count

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


class EffectSimulationClock:
    """Time<EffectSimulation> with its Real and Virtual parents (reference src/time.rs, bevy_time): the host-side
    producer of GpuSimParams. Method names follow EffectSimulationTime (time.rs:48-108)."""

    def __init__(self):
        self._h = lib.hnb_sim_clock_create()

    def __del__(self):
        try:
            if self._h:
                lib.hnb_sim_clock_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # EffectSimulationTime
    def relative_speed(self) -> float:
        return lib.hnb_sim_clock_relative_speed(self._h)

    def effective_speed(self) -> float:
        return lib.hnb_sim_clock_effective_speed(self._h)

    def set_relative_speed(self, ratio: float):
        check(lib.hnb_sim_clock_set_relative_speed(self._h, float(ratio)))

    def pause(self):
        lib.hnb_sim_clock_pause(self._h)

    def unpause(self):
        lib.hnb_sim_clock_unpause(self._h)

    def is_paused(self) -> bool:
        return bool(lib.hnb_sim_clock_is_paused(self._h))

    def was_paused(self) -> bool:
        return bool(lib.hnb_sim_clock_was_paused(self._h))

    # Time<Virtual>
    def set_virtual_relative_speed(self, ratio: float):
        check(lib.hnb_sim_clock_set_virtual_relative_speed(self._h, float(ratio)))

    def set_virtual_paused(self, paused: bool):
        lib.hnb_sim_clock_set_virtual_paused(self._h, int(paused))

    def set_max_delta_ns(self, ns: int):
        check(lib.hnb_sim_clock_set_max_delta_ns(self._h, int(ns)))

    def advance(self, real_delta_ns: int):
        """time_system + effect_simulation_time_system (time.rs:164-183) for one frame."""
        check(lib.hnb_sim_clock_advance(self._h, int(real_delta_ns)))

    def sim_params(self, num_effects: int = 0) -> "N.SimParams":
        """extract_sim_params + GpuSimParams::from (mod.rs:2796-2811, :266-279)."""
        out = N.SimParams()
        check(lib.hnb_sim_clock_sim_params(self._h, num_effects, C.byref(out)))
        return out

    @property
    def state(self) -> SimClockState:
        s = SimClockState()
        check(lib.hnb_sim_clock_state(self._h, C.byref(s)))
        return s


def default_sim_params() -> "N.SimParams":
    """GpuSimParams::default (mod.rs:244-256)."""
    out = N.SimParams()
    lib.hnb_sim_params_default(C.byref(out))
    return out
