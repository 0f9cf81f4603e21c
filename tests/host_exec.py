"""Runs GENERATED effect code on the CPU: test infrastructure for `-m "not gpu"` runs.

The translation unit the runtime hands to NVRTC is, up to the kernel templates, portable C++: the WGSL vocabulary
(`hnb_wgsl.cuh`), the table structs, the effect's `Particle` / pack / unpack code and the two generated bodies
`hnb_init_body` / `hnb_update_body`. This module cuts the TU before the kernel templates, adds a 20-line shim for the
few CUDA-isms it uses (`__device__`, `float4`, the `__float_as_uint` family) and a plain loop that calls the bodies for a list
of particles, and builds it with g++ (`-ffp-contract=off`, no fast-math: IEEE like the device build). Comparing its
results with the numpy interpreter checks the *values* the lowered text computes — operand order, hoisting, type
conversions — without a GPU; the GPU suite then only has to show that the device executes the same text the same way.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "build" / "host_exec"
MARK = "// hnb_particle_kernels.cuh"

PRELUDE = r"""
#include <math.h>
#include <stdint.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
"""

DRIVER = r"""
namespace hnb {
static void host_ctx(Ctx& c, const SimParams* sim, const Spawner* sp, const void* props, u32 particle_index, u32 particle_counter) {
    memset((void*)&c, 0, sizeof(c));
    c.particle_index = particle_index;
    c.particle_counter = particle_counter;
    c.seed = pcg_hash(particle_index ^ sp->seed);  // vfx_init.wgsl:154, vfx_update.wgsl:138
    c.is_alive = true;
    c.sim = sim;
    c.props = (const Properties*)props;
    c.spawner = sp;
    c.transform = hnb_transform_from_rows(sp->transform, sp->transform + 4, sp->transform + 8);
    c.inverse_transform = hnb_transform_from_rows(sp->inverse_transform, sp->inverse_transform + 4, sp->inverse_transform + 8);
}
}
// records: n AoS rows of `stride` bytes (the reference layout == the planes of RawParticle back to back)
extern "C" void host_update(uint8_t* records, uint32_t n, uint32_t stride, const uint32_t* particle_index, const void* sim, const void* spawner,
                            const void* props, uint8_t* is_alive_out) {
    for (uint32_t i = 0; i < n; ++i) {
        hnb::RawParticle raw;
        memset((void*)&raw, 0, sizeof(raw));
        memcpy((void*)&raw, records + (size_t)i * stride, stride);
        hnb::Particle particle;
        hnb::hnb_unpack(raw, particle);
        hnb::Ctx ctx;
        hnb::host_ctx(ctx, (const hnb::SimParams*)sim, (const hnb::Spawner*)spawner, props, particle_index[i], 0u);
        const bool alive = hnb::hnb_update_body(particle, ctx);
        hnb::hnb_pack<false>(particle, raw);
        memcpy(records + (size_t)i * stride, (const void*)&raw, stride);
        is_alive_out[i] = alive ? 1 : 0;
    }
}
extern "C" void host_init(uint8_t* records, uint32_t n, uint32_t stride, const uint32_t* particle_index, uint32_t first_particle_counter, const void* sim,
                          const void* spawner, const void* props) {
    for (uint32_t i = 0; i < n; ++i) {
        hnb::Ctx ctx;
        hnb::host_ctx(ctx, (const hnb::SimParams*)sim, (const hnb::Spawner*)spawner, props, particle_index[i], first_particle_counter + i);
        hnb::Particle particle = hnb::Particle();
        hnb::hnb_init_body(particle, ctx);
        hnb::RawParticle raw;
        memset((void*)&raw, 0, sizeof(raw));
        hnb::hnb_pack<true>(particle, raw);
        memcpy(records + (size_t)i * stride, (const void*)&raw, stride);
    }
}
"""


class HostEffect:
    """The generated bodies of one lowered effect, callable on numpy AoS records."""

    def __init__(self, lowered):
        src = lowered.generate_source()
        if "#define HNB_EMIT_EVENTS 1" in src or "#define HNB_READ_PARENT 1" in src:
            raise NotImplementedError("host execution covers effects without GPU spawn events")
        cut = src.index(MARK)
        text = PRELUDE + src[:cut] + DRIVER
        OUT.mkdir(parents=True, exist_ok=True)
        tag = hashlib.sha1(text.encode()).hexdigest()[:16]
        cpp, so = OUT / f"fx_{tag}.cpp", OUT / f"fx_{tag}.so"
        if not so.exists():
            cpp.write_text(text)
            cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", str(cpp), "-o", str(so)]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError("host build of the generated effect failed:\n" + proc.stderr[:4000])
        self.lib = C.CDLL(str(so))
        self.stride = lowered.particle_stride
        vp, u32 = C.c_void_p, C.c_uint32
        self.lib.host_update.argtypes = [vp, u32, u32, vp, vp, vp, vp, vp]
        self.lib.host_update.restype = None
        self.lib.host_init.argtypes = [vp, u32, u32, vp, u32, vp, vp, vp]
        self.lib.host_init.restype = None

    def update(self, records: np.ndarray, particle_index: np.ndarray, sim, spawner, props: bytes | None = None):
        """records: (n, stride/4) uint32, updated in place; returns the is_alive flags."""
        rec = np.ascontiguousarray(records, dtype=np.uint32)
        idx = np.ascontiguousarray(particle_index, dtype=np.uint32)
        alive = np.zeros(len(rec), dtype=np.uint8)
        pb = C.create_string_buffer(props, len(props)) if props else None
        self.lib.host_update(rec.ctypes.data, len(rec), self.stride, idx.ctypes.data, C.addressof(sim), C.addressof(spawner),
                             C.addressof(pb) if pb else None, alive.ctypes.data)
        records[:] = rec
        return alive.astype(bool)

    def init(self, n: int, particle_index: np.ndarray, first_particle_counter: int, sim, spawner, props: bytes | None = None) -> np.ndarray:
        rec = np.zeros((n, self.stride // 4), dtype=np.uint32)
        idx = np.ascontiguousarray(particle_index, dtype=np.uint32)
        pb = C.create_string_buffer(props, len(props)) if props else None
        self.lib.host_init(rec.ctypes.data, n, self.stride, idx.ctypes.data, first_particle_counter, C.addressof(sim), C.addressof(spawner),
                           C.addressof(pb) if pb else None)
        return rec


def replay_frame(host: HostEffect, oracle_effect, ref, orc, props_blob: bytes | None = None):
    """Advance `ref` by one frame with the numpy oracle and replay both passes with the generated code on the states
    the oracle saw. Returns (init_records_host, init_records_oracle, update_records_host, update_records_oracle,
    alive_host, alive_oracle) for single-instance worlds."""
    md, sp = ref.metadata[0], ref.spawners[0]
    base = sp.slab_offset
    alive0, counter0 = md.alive_count, md.particle_counter
    oracle_effect.init_pass(ref)
    spawned = md.alive_count - alive0
    col_w = md.indirect_write_index
    new_idx = ref.indirect[base + alive0: base + alive0 + spawned, col_w].copy()
    init_oracle = ref.particles[base + new_idx.astype(np.int64)].copy()
    init_host = host.init(spawned, new_idx, counter0, ref.sim, sp, props_blob)
    ref.oracle_indirect(orc)
    ref.oracle_prefix_sum(orc)
    n = md.max_update
    read_col = 1 - md.indirect_write_index
    upd_idx = ref.indirect[base: base + n, read_col].copy()
    before = ref.particles[base + upd_idx.astype(np.int64)].copy()
    oracle_effect.update_pass(ref)
    upd_oracle = ref.particles[base + upd_idx.astype(np.int64)].copy()
    upd_host = before.copy()
    alive_host = host.update(upd_host, upd_idx, ref.sim, sp, props_blob)
    survivors = set(ref.indirect[base: base + md.alive_count, md.indirect_write_index].tolist())
    alive_oracle = np.array([int(p) in survivors for p in upd_idx], dtype=bool)
    return init_host, init_oracle, upd_host, upd_oracle, alive_host, alive_oracle
