"""TEST INFRASTRUCTURE ONLY — CPU restatement of the simulation clock that produces GpuSimParams.

Only tests/ may import this module; the product (bevy_hanabi_b200/csrc/graph/sim_clock.cpp) never does.

Restates, with exact rational arithmetic where Rust works on integers:
  * EffectSimulation / EffectSimulationTime            reference src/time.rs:30-162
  * effect_simulation_time_system                      reference src/time.rs:164-183
  * extract_sim_params                                 reference src/render/mod.rs:2796-2811
  * From<&SimParams> for GpuSimParams, Default         reference src/render/mod.rs:244-279
  * Time<Real> / Time<Virtual> update and Duration::{mul_f64, from_secs_f64 (round to nearest, ties to even),
    as_secs_f32, as_secs_f64}: bevy_time 0.19 and Rust core (un-vendored dependencies, reference Cargo.toml:78,
    rust-version 1.95), published algorithm restated.

Parity unpinned beyond the reference's own test (time.rs:207-254), which pins the *relative* values of the three
clocks within 1e-6 and is replayed by tests/test_sim_clock_cpu.py.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

NS = 10**9
U64_MAX = 2**64 - 1
f32 = np.float32


def as_secs_f64(ns: int) -> float:
    """Duration::as_secs_f64: (secs as f64) + (nanos as f64) / 1e9."""
    return float(ns // NS) + float(ns % NS) / 1e9


def as_secs_f32(ns: int) -> np.float32:
    """Duration::as_secs_f32: (secs as f32) + (nanos as f32) / 1e9f32."""
    return f32(f32(ns // NS) + f32(ns % NS) / f32(1e9))


def from_secs_f64(secs: float) -> int:
    """Duration::from_secs_f64: nearest nanosecond to the exact value of the double, ties to even."""
    if math.isnan(secs) or math.isinf(secs) or secs < 0.0:
        raise OverflowError("value is either too big or NaN")
    exact = Fraction(secs) * NS          # Fraction(float) is exact
    q, r = divmod(exact.numerator, exact.denominator)
    twice = 2 * r
    if twice > exact.denominator or (twice == exact.denominator and (q & 1)):
        q += 1
    if q > U64_MAX:
        raise OverflowError("value is either too big or NaN")
    return q


def mul_f64(ns: int, x: float) -> int:
    """Duration::mul_f64."""
    return from_secs_f64(x * as_secs_f64(ns))


class _Time:
    def __init__(self):
        self.delta = 0
        self.elapsed = 0

    def advance_by(self, d: int):
        self.delta = d
        self.elapsed += d


class SimClockOracle:
    def __init__(self):
        self.real, self.virt, self.sim = _Time(), _Time(), _Time()
        self.max_delta = 250_000_000
        self.virt_paused = False
        self.virt_relative_speed = 1.0
        self.virt_effective_speed = 1.0
        self.paused = False                 # EffectSimulation::default, time.rs:37-45
        self.relative_speed = 1.0
        self.effective_speed = 1.0

    def set_relative_speed(self, ratio: float):  # time.rs:137-141
        if not math.isfinite(ratio):
            raise ValueError("tried to go infinitely fast")
        if not ratio >= 0.0:
            raise ValueError("tried to go back in time")
        self.relative_speed = ratio

    def was_paused(self) -> bool:  # time.rs:159
        return self.effective_speed == 0.0

    def advance(self, raw_ns: int):
        clamped = min(raw_ns, self.max_delta)
        v_speed = 0.0 if self.virt_paused else self.virt_relative_speed
        v_delta = mul_f64(clamped, v_speed) if v_speed != 1.0 else clamped
        s_speed = 0.0 if self.paused else self.relative_speed              # time.rs:169-173
        s_delta = mul_f64(v_delta, s_speed) if s_speed != 1.0 else v_delta  # time.rs:174-179
        self.real.advance_by(raw_ns)
        self.virt_effective_speed = v_speed
        self.virt.advance_by(v_delta)
        self.effective_speed = s_speed * v_speed                           # time.rs:181
        self.sim.advance_by(s_delta)                                       # time.rs:182

    def gpu_sim_params(self, num_effects: int = 0) -> dict:
        """mod.rs:2806-2811 then mod.rs:266-279; field order of GpuSimParams (mod.rs:218-242)."""
        return {
            "delta_time": as_secs_f32(self.sim.delta),
            "time": f32(as_secs_f64(self.sim.elapsed)),
            "virtual_delta_time": as_secs_f32(self.virt.delta),
            "virtual_time": f32(as_secs_f64(self.virt.elapsed)),
            "real_delta_time": as_secs_f32(self.real.delta),
            "real_time": f32(as_secs_f64(self.real.elapsed)),
            "num_effects": num_effects,
        }


def default_gpu_sim_params() -> dict:
    """GpuSimParams::default, mod.rs:244-256."""
    return {"delta_time": f32(0.04), "time": f32(0.0), "virtual_delta_time": f32(0.04), "virtual_time": f32(0.0),
            "real_delta_time": f32(0.04), "real_time": f32(0.0), "num_effects": 0}
