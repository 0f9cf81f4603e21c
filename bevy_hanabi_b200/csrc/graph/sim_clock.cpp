// sim_clock.cpp — the CPU producer of GpuSimParams, the one per-frame table every kernel of the path reads
// (caller-side step before the path, §8f-3):
//   * EffectSimulation / EffectSimulationTime            (reference src/time.rs:30-162)
//   * effect_simulation_time_system                      (reference src/time.rs:164-183)
//   * SimParams, extract_sim_params                      (reference src/render/mod.rs:193-212, :2796-2811)
//   * From<&SimParams> for GpuSimParams, Default         (reference src/render/mod.rs:244-279)
// The Real and Virtual clocks the reference reads belong to bevy_time 0.19 (un-vendored dependency, Cargo.toml:78);
// their published update rule is restated here on integer nanoseconds (Rust `Duration`):
//   Time<Real>:    delta = raw; elapsed += delta
//   Time<Virtual>: delta = min(raw, max_delta) * (paused ? 0 : relative_speed), scaled only when speed != 1.0
//   Duration::mul_f64(x)   = from_secs_f64(x * as_secs_f64());  from_secs_f64 rounds to nearest, ties to even
//   delta_secs (f32)       = secs as f32 + nanos as f32 / 1e9f;  elapsed_secs_f64 = secs as f64 + nanos as f64 / 1e9
// "parity unpinned" beyond the reference's own test (time.rs:207-254: relative values of the three clocks within
// 1e-6), which tests/test_sim_clock_cpu.py replays with injected real deltas instead of sleeps.
// Pure host code, exposed through include/hanabi_b200_graph.h.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "hanabi_b200_graph.h"

extern "C" void hnb_set_last_error_(const char* msg);

namespace {

constexpr uint64_t kNsPerSec = 1000000000ull;
typedef unsigned __int128 u128;

struct Clock {  // bevy_time Time<T>: delta and elapsed as Durations
    uint64_t delta_ns = 0;
    u128 elapsed_ns = 0;
    void advance_by(uint64_t d) {
        delta_ns = d;
        elapsed_ns += d;
    }
    float delta_secs() const {  // Duration::as_secs_f32
        return (float)(delta_ns / kNsPerSec) + (float)(uint32_t)(delta_ns % kNsPerSec) / 1000000000.0f;
    }
    double elapsed_secs_f64() const {  // Duration::as_secs_f64
        return (double)(uint64_t)(elapsed_ns / kNsPerSec) + (double)(uint32_t)(elapsed_ns % kNsPerSec) / 1000000000.0;
    }
};

double as_secs_f64(uint64_t ns) { return (double)(ns / kNsPerSec) + (double)(uint32_t)(ns % kNsPerSec) / 1000000000.0; }

// Duration::from_secs_f64: the exact binary value of `secs`, in nanoseconds, rounded to nearest (ties to even).
// false = negative, non-finite or beyond u64 nanoseconds (Rust: panics "value is either too big or NaN").
bool from_secs_f64(double secs, uint64_t* out_ns) {
    if (!(secs >= 0.0) || !std::isfinite(secs)) return false;
    if (secs == 0.0) { *out_ns = 0; return true; }
    int e = 0;
    const double fr = std::frexp(secs, &e);           // secs = fr * 2^e, fr in [0.5, 1)
    const uint64_t m = (uint64_t)std::ldexp(fr, 53);  // 53-bit integer mantissa, exact
    e -= 53;                                          // secs = m * 2^e
    const u128 prod = (u128)m * kNsPerSec;            // < 2^83
    u128 ns;
    if (e >= 0) {
        if (e > 44) return false;                     // >= 2^127
        ns = prod << e;
    } else {
        const int s = -e;
        if (s >= 127) { *out_ns = 0; return true; }
        const u128 q = prod >> s;
        const u128 rem = prod - (q << s);
        const u128 half = (u128)1 << (s - 1);
        ns = q + ((rem > half || (rem == half && (q & 1))) ? 1 : 0);
    }
    if (ns > (u128)UINT64_MAX) return false;
    *out_ns = (uint64_t)ns;
    return true;
}

bool mul_f64(uint64_t ns, double x, uint64_t* out) { return from_secs_f64(x * as_secs_f64(ns), out); }

int32_t bad(const char* msg) {
    hnb_set_last_error_(msg);
    return HNB_ERR_INVALID_ARG;
}

}  // namespace

struct hnb_sim_clock {
    Clock real, virt, sim;
    // Time<Virtual> context (bevy_time): max_delta 250 ms, relative speed 1
    uint64_t max_delta_ns = 250000000ull;
    bool virt_paused = false;
    double virt_relative_speed = 1.0, virt_effective_speed = 1.0;
    // EffectSimulation (time.rs:31-46)
    bool paused = false;
    double relative_speed = 1.0, effective_speed = 1.0;
};

extern "C" {

hnb_sim_clock* hnb_sim_clock_create(void) { return new hnb_sim_clock(); }
void hnb_sim_clock_destroy(hnb_sim_clock* c) { delete c; }

// Time<Virtual>::set_relative_speed_f64 / pause / unpause / set_max_delta (bevy_time): same assertions as below
int32_t hnb_sim_clock_set_virtual_relative_speed(hnb_sim_clock* c, double ratio) {
    if (!c) return bad("clock is NULL");
    if (!std::isfinite(ratio)) return bad("tried to go infinitely fast");
    if (!(ratio >= 0.0)) return bad("tried to go back in time");
    c->virt_relative_speed = ratio;
    return HNB_OK;
}
void hnb_sim_clock_set_virtual_paused(hnb_sim_clock* c, uint32_t paused) { if (c) c->virt_paused = paused != 0; }
int32_t hnb_sim_clock_set_max_delta_ns(hnb_sim_clock* c, uint64_t ns) {
    if (!c) return bad("clock is NULL");
    if (ns == 0) return bad("tried to set max delta to zero");
    c->max_delta_ns = ns;
    return HNB_OK;
}

// EffectSimulationTime::set_relative_speed_f64 (time.rs:137-141): the reference asserts, the C ABI reports
int32_t hnb_sim_clock_set_relative_speed(hnb_sim_clock* c, double ratio) {
    if (!c) return bad("clock is NULL");
    if (!std::isfinite(ratio)) return bad("tried to go infinitely fast");
    if (!(ratio >= 0.0)) return bad("tried to go back in time");
    c->relative_speed = ratio;
    return HNB_OK;
}
void hnb_sim_clock_pause(hnb_sim_clock* c) { if (c) c->paused = true; }     // time.rs:144
void hnb_sim_clock_unpause(hnb_sim_clock* c) { if (c) c->paused = false; }  // time.rs:149
uint32_t hnb_sim_clock_is_paused(const hnb_sim_clock* c) { return c && c->paused ? 1u : 0u; }                   // time.rs:154
uint32_t hnb_sim_clock_was_paused(const hnb_sim_clock* c) { return c && c->effective_speed == 0.0 ? 1u : 0u; }  // time.rs:159
double hnb_sim_clock_relative_speed(const hnb_sim_clock* c) { return c ? c->relative_speed : 0.0; }                  // time.rs:117
double hnb_sim_clock_effective_speed(const hnb_sim_clock* c) { return c ? c->effective_speed : 0.0; }                // time.rs:127

// One frame: bevy's time_system (Real, then Virtual from the real delta) followed by
// effect_simulation_time_system (time.rs:164-183). Nothing is modified when a product overflows.
int32_t hnb_sim_clock_advance(hnb_sim_clock* c, uint64_t real_delta_ns) {
    if (!c) return bad("clock is NULL");
    const uint64_t clamped = real_delta_ns > c->max_delta_ns ? c->max_delta_ns : real_delta_ns;
    const double v_speed = c->virt_paused ? 0.0 : c->virt_relative_speed;
    uint64_t v_delta = clamped;
    if (v_speed != 1.0 && !mul_f64(clamped, v_speed, &v_delta)) return bad("virtual delta overflows a Duration");
    const double s_speed = c->paused ? 0.0 : c->relative_speed;   // time.rs:169-173
    uint64_t s_delta = v_delta;                                    // "avoid rounding when at normal speed" (time.rs:177)
    if (s_speed != 1.0 && !mul_f64(v_delta, s_speed, &s_delta)) return bad("simulation delta overflows a Duration");
    c->real.advance_by(real_delta_ns);
    c->virt_effective_speed = v_speed;
    c->virt.advance_by(v_delta);
    c->effective_speed = s_speed * v_speed;                        // time.rs:181
    c->sim.advance_by(s_delta);
    return HNB_OK;
}

// extract_sim_params (mod.rs:2796-2811) then From<&SimParams> for GpuSimParams (mod.rs:266-279): times are kept in
// f64 and narrowed once; `num_effects` is filled in by prepare_effects (mod.rs:4471-4472).
int32_t hnb_sim_clock_sim_params(const hnb_sim_clock* c, uint32_t num_effects, hnb_sim_params* out) {
    if (!c || !out) return bad("clock or out is NULL");
    out->delta_time = c->sim.delta_secs();
    out->time = (float)c->sim.elapsed_secs_f64();
    out->virtual_delta_time = c->virt.delta_secs();
    out->virtual_time = (float)c->virt.elapsed_secs_f64();
    out->real_delta_time = c->real.delta_secs();
    out->real_time = (float)c->real.elapsed_secs_f64();
    out->num_effects = num_effects;
    return HNB_OK;
}

// GpuSimParams::default (mod.rs:244-256)
void hnb_sim_params_default(hnb_sim_params* out) {
    if (!out) return;
    out->delta_time = 0.04f;
    out->time = 0.f;
    out->virtual_delta_time = 0.04f;
    out->virtual_time = 0.f;
    out->real_delta_time = 0.04f;
    out->real_time = 0.f;
    out->num_effects = 0;
}

int32_t hnb_sim_clock_state(const hnb_sim_clock* c, hnb_sim_clock_state_t* out) {
    if (!c || !out) return bad("clock or out is NULL");
    const u128 cap = (u128)UINT64_MAX;
    out->real_elapsed_ns = (uint64_t)(c->real.elapsed_ns > cap ? cap : c->real.elapsed_ns);
    out->virtual_elapsed_ns = (uint64_t)(c->virt.elapsed_ns > cap ? cap : c->virt.elapsed_ns);
    out->sim_elapsed_ns = (uint64_t)(c->sim.elapsed_ns > cap ? cap : c->sim.elapsed_ns);
    out->real_delta_ns = c->real.delta_ns;
    out->virtual_delta_ns = c->virt.delta_ns;
    out->sim_delta_ns = c->sim.delta_ns;
    out->virtual_effective_speed = c->virt_effective_speed;
    out->sim_effective_speed = c->effective_speed;
    return HNB_OK;
}

}  // extern "C"
