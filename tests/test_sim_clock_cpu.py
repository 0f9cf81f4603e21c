"""The simulation clock that produces GpuSimParams (csrc/graph/sim_clock.cpp) against its oracle
(oracle/sim_clock_oracle.py) and the reference's own test (src/time.rs:207-254, replayed with injected real
deltas instead of sleeps). Integer nanoseconds and f32/f64 words are compared bit for bit."""
import math
import random
import struct

import numpy as np
import pytest

from bevy_hanabi_b200._native import HanabiError
from bevy_hanabi_b200.spawn import EffectSimulationClock, default_sim_params
from oracle.sim_clock_oracle import SimClockOracle, default_gpu_sim_params, from_secs_f64, mul_f64

FIELDS = ("delta_time", "time", "virtual_delta_time", "virtual_time", "real_delta_time", "real_time")


def _bits(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def _same(clock: EffectSimulationClock, oracle: SimClockOracle, num_effects=0):
    s = clock.state
    assert (s.real_elapsed_ns, s.real_delta_ns) == (oracle.real.elapsed, oracle.real.delta)
    assert (s.virtual_elapsed_ns, s.virtual_delta_ns) == (oracle.virt.elapsed, oracle.virt.delta)
    assert (s.sim_elapsed_ns, s.sim_delta_ns) == (oracle.sim.elapsed, oracle.sim.delta)
    assert s.virtual_effective_speed == oracle.virt_effective_speed
    assert s.sim_effective_speed == oracle.effective_speed
    got, want = clock.sim_params(num_effects), oracle.gpu_sim_params(num_effects)
    for f in FIELDS:
        assert _bits(getattr(got, f)) == _bits(want[f]), f
    assert got.num_effects == want["num_effects"]


def test_effect_simulation_time():
    """time.rs:207-254: the relations between the real, virtual and simulation clocks, within its EPSILON."""
    EPS = 0.000001
    c = EffectSimulationClock()
    c.advance(0)                        # first app.update(): no time has passed
    c.advance(1_234_567)                # "sleep 1 ms", default speeds
    p = c.sim_params()
    assert abs(p.virtual_delta_time - p.real_delta_time) < EPS
    assert abs(p.delta_time - p.real_delta_time) < EPS
    c.set_virtual_relative_speed(2.0)   # virtual speed 2.0
    c.advance(1_111_111)
    p = c.sim_params()
    assert abs(p.virtual_delta_time - 2.0 * p.real_delta_time) < EPS
    assert abs(p.delta_time - 2.0 * p.real_delta_time) < EPS
    assert abs(c.state.virtual_effective_speed - 2.0) < EPS
    assert abs(c.effective_speed() - 2.0) < EPS
    c.set_relative_speed(3.0)           # virtual speed 2.0 and effect speed 3.0
    c.advance(1_313_131)
    p = c.sim_params()
    assert abs(p.virtual_delta_time - 2.0 * p.real_delta_time) < EPS
    assert abs(p.delta_time - 6.0 * p.real_delta_time) < EPS
    assert abs(c.state.virtual_effective_speed - 2.0) < EPS
    assert abs(c.effective_speed() - 6.0) < EPS
    assert c.relative_speed() == 3.0


def test_defaults():
    c = EffectSimulationClock()       # EffectSimulation::default (time.rs:37-45)
    assert c.relative_speed() == 1.0 and c.effective_speed() == 1.0
    assert not c.is_paused() and not c.was_paused()
    d, want = default_sim_params(), default_gpu_sim_params()   # GpuSimParams::default (mod.rs:244-256)
    for f in FIELDS:
        assert _bits(getattr(d, f)) == _bits(want[f])
    assert d.num_effects == 0
    _same(c, SimClockOracle())        # before the first frame every clock reads zero


def test_pause_and_was_paused():
    c, o = EffectSimulationClock(), SimClockOracle()
    for x in (c, o):
        x.advance(16_666_667)
    c.pause(); o.paused = True
    assert c.is_paused() and not c.was_paused()          # was_paused reflects the last *advanced* frame (time.rs:159)
    for x in (c, o):
        x.advance(16_666_667)
    assert c.was_paused() and o.was_paused()
    p = c.sim_params()
    assert p.delta_time == 0.0 and p.virtual_delta_time > 0.0
    _same(c, o)
    c.unpause(); o.paused = False
    for x in (c, o):
        x.advance(16_666_667)
    assert not c.was_paused() and c.sim_params().delta_time > 0.0
    _same(c, o)
    # a paused virtual clock stops the simulation clock too, and its effective speed reads 0
    c.set_virtual_paused(True); o.virt_paused = True
    for x in (c, o):
        x.advance(16_666_667)
    assert c.was_paused() and c.sim_params().virtual_delta_time == 0.0 and c.sim_params().real_delta_time > 0.0
    _same(c, o)


def test_speed_validation_matches_the_reference_assertions():
    c = EffectSimulationClock()
    for bad, msg in ((math.inf, "infinitely fast"), (math.nan, "infinitely fast"), (-1.0, "back in time")):
        with pytest.raises(HanabiError, match=msg):       # time.rs:138-139
            c.set_relative_speed(bad)
        with pytest.raises(HanabiError, match=msg):
            c.set_virtual_relative_speed(bad)
    assert c.relative_speed() == 1.0                       # nothing changed
    c.set_relative_speed(0.0)                              # zero is allowed: the clock stands still
    c.advance(5_000_000)
    assert c.sim_params().delta_time == 0.0 and c.was_paused()
    with pytest.raises(HanabiError):
        c.set_max_delta_ns(0)


def test_max_delta_clamps_the_virtual_clock_only():
    c, o = EffectSimulationClock(), SimClockOracle()
    for x in (c, o):
        x.advance(2 * 10**9)                               # a 2 s hitch: virtual time moves 250 ms
    s = c.state
    assert s.real_delta_ns == 2 * 10**9 and s.virtual_delta_ns == 250_000_000 and s.sim_delta_ns == 250_000_000
    _same(c, o)
    c.set_max_delta_ns(10**9); o.max_delta = 10**9
    for x in (c, o):
        x.advance(2 * 10**9)
    assert c.state.virtual_delta_ns == 10**9
    _same(c, o)


def test_normal_speed_takes_the_unrounded_delta():
    """time.rs:174-179: at speed 1.0 the delta is handed on as is (no float round trip)."""
    c = EffectSimulationClock()
    c.advance(123_456_789)
    s = c.state
    assert s.virtual_delta_ns == s.sim_delta_ns == 123_456_789


def test_from_secs_f64_rounds_to_nearest_even():
    assert from_secs_f64(0.999e-9) == 1 and from_secs_f64(0.4e-9) == 0
    assert from_secs_f64(2.5e-9) in (2, 3)                 # 2.5e-9 is not exactly representable: nearest of the double
    assert from_secs_f64(0.5 ** 31) == 0                   # 0.4656...ns
    assert from_secs_f64(1.5) == 1_500_000_000
    assert mul_f64(16_666_667, 3.0) == 50_000_001
    with pytest.raises(OverflowError):
        from_secs_f64(-0.0 - 1e-30)


@pytest.mark.parametrize("seed", range(6))
def test_random_sessions_equal_the_oracle(seed):
    rng = random.Random(1000 + seed)
    c, o = EffectSimulationClock(), SimClockOracle()
    for frame in range(400):
        r = rng.random()
        if r < 0.08:
            v = rng.choice([0.0, 0.25, 0.5, 1.0, 1.0, 2.0, 3.0, rng.uniform(0.0, 8.0), rng.uniform(0.0, 1e-3)])
            c.set_relative_speed(v); o.set_relative_speed(v)
        elif r < 0.14:
            v = rng.choice([0.0, 0.5, 1.0, 1.0, 2.0, rng.uniform(0.0, 4.0)])
            c.set_virtual_relative_speed(v); o.virt_relative_speed = v
        elif r < 0.18:
            c.pause(); o.paused = True
        elif r < 0.24:
            c.unpause(); o.paused = False
        elif r < 0.27:
            b = rng.random() < 0.5
            c.set_virtual_paused(b); o.virt_paused = b
        elif r < 0.29:
            m = rng.choice([1, 1000, 16_666_667, 250_000_000, 10**9])
            c.set_max_delta_ns(m); o.max_delta = m
        raw = rng.choice([0, 1, 999, rng.randrange(0, 40_000_000), rng.randrange(0, 40_000_000), 16_666_667, 8_333_333,
                          rng.randrange(0, 3 * 10**9)])
        c.advance(raw); o.advance(raw)
        assert c.is_paused() == o.paused and c.was_paused() == o.was_paused()
        assert c.relative_speed() == o.relative_speed
        _same(c, o, num_effects=frame)


def test_long_run_keeps_f64_time():
    """SimParams keeps elapsed times in f64 and narrows once per frame (mod.rs:197-209, :269-274): after a day of
    frames the f32 `time` still equals the f64 sum rounded once, not an f32 running sum."""
    c, o = EffectSimulationClock(), SimClockOracle()
    frames, dt = 5_184_000, 16_666_667           # 24 h at 60 Hz; one call per 43 200 frames to keep the test fast
    step = 43_200
    c.set_max_delta_ns(2**62); o.max_delta = 2**62
    for _ in range(frames // step):
        c.advance(dt * step); o.advance(dt * step)
    _same(c, o)
    t = c.sim_params().time
    assert t == np.float32(frames * dt / 1e9)
    running = np.float32(0.0)
    for _ in range(2000):
        running = np.float32(running + np.float32(dt / 1e9))
    assert running != np.float32(2000 * dt / 1e9)   # what the f64 bookkeeping avoids
