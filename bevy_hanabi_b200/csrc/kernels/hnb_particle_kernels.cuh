// hnb_particle_kernels.cuh — hand-written sm_100a kernel templates for the two per-particle passes
// of the hot path. They play the role of the reference's WGSL templates
//   src/render/vfx_init.wgsl   (entry :101-196)   → hnb_init
//   src/render/vfx_update.wgsl (entry :106-167)   → hnb_update
// and are specialised per effect by textual inclusion after a generated section (see
// runtime/effect_source.cpp) which must define, inside namespace hnb:
//
//   HNB_NUM_PLANES, HNB_HAS_PROPERTIES, HNB_CONSUME_EVENTS, HNB_EMIT_EVENTS, HNB_READ_PARENT,
//   HNB_RELAXED_ORDER, HNB_TILE_K
//   struct Particle;  struct Properties;  struct RawParticle;  [struct ParentParticle]
//   hnb_load_raw / hnb_store_raw / hnb_unpack / hnb_pack       (SoA planes <-> Particle)
//   hnb_init_body(Particle&, Ctx&)            = {{INIT_CODE}} + PREV/NEXT reset + {{SIM_SPACE}}
//   hnb_update_body(Particle&, Ctx&) -> bool  = {{AGE_CODE}} {{REAP_CODE}} {{UPDATE_CODE}}, returns is_alive
//
// Design:
//   * update is a persistent, single-pass "process + stable compaction" kernel. A CTA takes tiles of
//     HNB_TILE rows of ONE effect instance from a ticket counter; rows are read through the alive list
//     (coalesced u32), particles through float4 SoA planes, processed in registers, written back, and
//     the survivors' indices are compacted into the write list in row order using warp ballots inside
//     the tile and a decoupled look-back chain across the tiles of the same instance. Dead rows are
//     pushed on the dead stack in the same canonical (row) order. The only atomic on the path is the
//     tile ticket: the per-particle contended atomics of the reference (vfx_update.wgsl:150,160,164)
//     are replaced by exact ranks, which also makes list ORDER deterministic (= serial thread order).
//   * init pops dead slots by rank as well: thread k of an instance takes dead[alive_count + k]
//     (vfx_init.wgsl:141-143 in serial order); the alive_count / particle_counter increments are
//     applied by the bookkeeping kernel that follows (hnb_static_kernels.cu).
#pragma once

namespace hnb {

#define HNB_BLOCK 256
#define HNB_WARPS (HNB_BLOCK / 32)
#define HNB_TILE (HNB_BLOCK * HNB_TILE_K)
#ifndef HNB_SMEM_EFFECTS
#define HNB_SMEM_EFFECTS 2048  // tile_prefix entries staged in shared memory (8 KB)
#endif
#ifndef HNB_MIN_BLOCKS
#define HNB_MIN_BLOCKS 3
#endif

// --- tile state word of the decoupled look-back: [63:34] epoch | [33:32] flag | [31:0] value ---
#define HNB_FLAG_AGGREGATE 1ull
#define HNB_FLAG_PREFIX 2ull
HNB_DI u64 hnb_pack_state(u32 epoch, u64 flag, u32 value) { return (u64(epoch) << 34) | (flag << 32) | u64(value); }
HNB_DI void hnb_st_state(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
HNB_DI u64 hnb_ld_state(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

HNB_DI u32 hnb_lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// find_location_from_particle (vfx_update.wgsl:51-72 / vfx_init.wgsl:51-72): upper bound of `x` in
// prefix[lo,hi), minus one. `prefix[lo]` is 0 by construction so the result is always >= lo.
template <typename Ptr> HNB_DI u32 hnb_find_effect(Ptr prefix, u32 lo, u32 hi, u32 x) {
    while (lo < hi) {
        const u32 mid = (hi + lo) >> 1u;
        if (x >= prefix[mid]) lo = mid + 1u; else hi = mid;
    }
    return lo - 1u;
}

// ---------------------------------------------------------------------------------------------
// init  ≙ vfx_init.wgsl main()
// ---------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(HNB_BLOCK) hnb_init(const BatchParams P) {
    const u32 thread_index = blockIdx.x * HNB_BLOCK + threadIdx.x;  // global_invocation_id.x
    if (thread_index >= P.init_thread_count) return;

    const BatchInfo bi = *P.batch_info;
    // Location in the packed init space of this batch (CPU prefix sums of spawn counts, batch.rs:358-383)
    const u32 slot = hnb_find_effect(P.spawn_prefix, bi.prefix_sum_offset, bi.prefix_sum_offset + bi.prefix_sum_count, thread_index);
    const u32 effect_index = slot - bi.prefix_sum_offset;
    const u32 update_index = thread_index - P.spawn_prefix[slot];
    const Spawner* spawner = &P.spawners[bi.spawner_base + effect_index];
    const u32 base_particle = spawner->slab_offset;
    const EffectMetadata* md = &P.metadata[spawner->effect_metadata_index];

    // Cap to the number of dead particles (vfx_init.wgsl:115-119)
    const u32 max_spawn = md->max_spawn;
    if (update_index >= max_spawn) return;
#if HNB_CONSUME_EVENTS
    const u32 event_index = update_index;
    const i32 event_count = P.child_infos[md->global_child_index].event_count;
    if (event_index >= u32(event_count)) return;
#else
    const u32 spawn_count = u32(spawner->spawn);
    if (update_index >= spawn_count) return;
#endif

    // Recycle a dead slot. Serial-order equivalent of `atomicAdd(alive_count, 1)` (:141): every thread
    // with a smaller update_index also passed the caps above, so this thread's rank IS update_index.
    const u32 alive_index = md->alive_count + update_index;
    const u32 slab_particle_dead_index = P.slab.dead_index[base_particle + alive_index];
    const u32 particle_index = slab_particle_dead_index - base_particle;

    Ctx hnb_ctx;
    hnb_ctx.particle_index = particle_index;
    hnb_ctx.particle_counter = md->particle_counter + update_index;  // atomicAdd(particle_counter, 1) (:151)
    hnb_ctx.seed = pcg_hash(particle_index ^ spawner->seed);         // :154
    hnb_ctx.sim = &P.frame->sim;
    hnb_ctx.spawner = spawner;
    hnb_ctx.transform = hnb_transform_from_rows(spawner->transform, spawner->transform + 4, spawner->transform + 8);
    hnb_ctx.inverse_transform = hnb_transform_from_rows(spawner->inverse_transform, spawner->inverse_transform + 4,
                                                        spawner->inverse_transform + 8);
    hnb_ctx.is_alive = true;
#if HNB_HAS_PROPERTIES
    hnb_ctx.props = (const Properties*)((const char*)P.properties + size_t(md->properties_array_index) * P.properties_stride);
#else
    hnb_ctx.props = nullptr;
#endif
#if HNB_READ_PARENT
    {
        const u32 parent_base_particle = spawner->parent_slab_offset;
        hnb_ctx.parent_particle_index = P.consume_events[event_index];
        ParentRawParticle praw;
        hnb_parent_load_raw(praw, P.parent_slab, parent_base_particle + hnb_ctx.parent_particle_index);
        hnb_parent_unpack(praw, hnb_ctx.parent_particle);
    }
#endif

    Particle particle = Particle();
    hnb_init_body(particle, hnb_ctx);

    // Append to the alive list (:191-192) and write the particle back (:195)
    const u32 write_index = md->indirect_write_index;
    P.slab.particle_index[write_index][base_particle + alive_index] = particle_index;
    RawParticle raw;
    hnb_raw_zero(raw);
    hnb_pack<true>(particle, raw);  // init also stores PREV/NEXT (vfx_init.wgsl:175-181)
    hnb_store_raw(raw, P.slab, base_particle + particle_index);
}

// ---------------------------------------------------------------------------------------------
// update  ≙ vfx_update.wgsl main()
// ---------------------------------------------------------------------------------------------
struct UpdateShared {
    u32 tile;
    u32 alive_before;
    u32 tile_alive;
    u32 warp_alive[HNB_TILE_K * HNB_WARPS];  // per (k, warp) alive counts, then their exclusive scan
    u32 tile_prefix[HNB_SMEM_EFFECTS + 1];
};

extern "C" __global__ void __launch_bounds__(HNB_BLOCK, HNB_MIN_BLOCKS) hnb_update(const BatchParams P) {
    __shared__ UpdateShared sh;
#if HNB_HAS_PROPERTIES
    __shared__ __align__(16) unsigned char sh_props[sizeof(Properties)];
#endif
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 warp = tid >> 5u;

    const BatchInfo bi = *P.batch_info;
    const u32 n_effects = bi.prefix_sum_count;
    const u32* g_tile_prefix = P.tile_prefix + bi.prefix_sum_offset;
    const u32 total_tiles = *P.batch_tiles;
    const u32 epoch = P.frame->epoch;
    const bool staged = n_effects <= HNB_SMEM_EFFECTS;
    if (staged) {
        for (u32 i = tid; i < n_effects; i += HNB_BLOCK) sh.tile_prefix[i] = g_tile_prefix[i];
    }
    u32 staged_effect = HNB_INVALID;
    (void)staged_effect;

    for (;;) {
        __syncthreads();  // previous tile fully done (protects sh.* and sh_props); also covers the staging above
        if (tid == 0) sh.tile = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const u32 tile = sh.tile;
        if (tile >= total_tiles) break;

        // Which instance does this tile belong to? (per-CTA replacement of the per-thread binary search
        // of vfx_update.wgsl:51-72; all threads read the same shared words: broadcast, no conflicts)
        u32 effect_index, tile_in_effect;
        if (staged) {
            effect_index = hnb_find_effect(sh.tile_prefix, 0u, n_effects, tile);
            tile_in_effect = tile - sh.tile_prefix[effect_index];
        } else {
            effect_index = hnb_find_effect(g_tile_prefix, 0u, n_effects, tile);
            tile_in_effect = tile - g_tile_prefix[effect_index];
        }
        Spawner* spawner = &P.spawners[bi.spawner_base + effect_index];
        const u32 base_particle = spawner->slab_offset;
        const u32 spawner_seed = spawner->seed;
        EffectMetadata* md = &P.metadata[spawner->effect_metadata_index];
        const u32 max_update = md->max_update;  // :119
        const u32 write_index = md->indirect_write_index;
        const u32 read_index = 1u - write_index;
        const u32* __restrict__ read_col = P.slab.particle_index[read_index] + base_particle;
        u32* __restrict__ write_col = P.slab.particle_index[write_index] + base_particle;

        Ctx hnb_ctx;
        hnb_ctx.sim = &P.frame->sim;
        hnb_ctx.spawner = spawner;
        hnb_ctx.particle_counter = 0u;
#if HNB_HAS_PROPERTIES
        if (staged_effect != effect_index) {
            const u32* src = (const u32*)((const char*)P.properties + size_t(md->properties_array_index) * P.properties_stride);
            for (u32 i = tid; i < sizeof(Properties) / 4u; i += HNB_BLOCK) ((u32*)sh_props)[i] = src[i];
            staged_effect = effect_index;
            __syncthreads();
        }
        hnb_ctx.props = (const Properties*)sh_props;
#else
        hnb_ctx.props = nullptr;
#endif
#if HNB_EMIT_EVENTS
        hnb_ctx.child_infos = P.child_infos;
        hnb_ctx.base_child_index = md->base_child_index;
        for (int i = 0; i < HNB_MAX_EVENT_BINDINGS; ++i) {
            hnb_ctx.emit_events[i] = P.emit_events[i];
            hnb_ctx.emit_events_capacity[i] = P.emit_events_capacity[i];
        }
#endif

        const u32 row0 = tile_in_effect * HNB_TILE;

        // 1) alive-list entries of this tile: row = row0 + k*BLOCK + tid (coalesced)
        u32 pidx[HNB_TILE_K];
        bool valid[HNB_TILE_K];
#pragma unroll
        for (int k = 0; k < HNB_TILE_K; ++k) {
            const u32 row = row0 + k * HNB_BLOCK + tid;
            valid[k] = row < max_update;
            pidx[k] = valid[k] ? read_col[row] : 0u;
        }
        // 2) gather the particle records (all loads in flight before any use)
        RawParticle raw[HNB_TILE_K];
#pragma unroll
        for (int k = 0; k < HNB_TILE_K; ++k) {
            if (valid[k]) hnb_load_raw(raw[k], P.slab, base_particle + pidx[k]);
            else hnb_raw_zero(raw[k]);
        }
        // 3) simulate + write back (WRITEBACK_CODE: every attribute except PREV/NEXT, lib.rs:1270-1281)
        bool alive[HNB_TILE_K];
#pragma unroll
        for (int k = 0; k < HNB_TILE_K; ++k) {
            alive[k] = false;
            if (valid[k]) {
                Particle particle;
                hnb_unpack(raw[k], particle);
                hnb_ctx.particle_index = pidx[k];
                hnb_ctx.seed = pcg_hash(pidx[k] ^ spawner_seed);  // :138
                hnb_ctx.is_alive = true;
                alive[k] = hnb_update_body(particle, hnb_ctx);
                hnb_pack<false>(particle, raw[k]);
                hnb_store_raw(raw[k], P.slab, base_particle + pidx[k]);
            }
        }
        // 4) ranks of the survivors inside the tile (row order = (k, warp, lane))
        u32 rank[HNB_TILE_K];
#pragma unroll
        for (int k = 0; k < HNB_TILE_K; ++k) {
            const u32 ballot = __ballot_sync(0xffffffffu, alive[k]);
            rank[k] = __popc(ballot & hnb_lanemask_lt());
            if (lane == 0) sh.warp_alive[k * HNB_WARPS + warp] = __popc(ballot);
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive scan of the K*WARPS (<= 32) counts
            const u32 n = HNB_TILE_K * HNB_WARPS;
            const u32 v = lane < n ? sh.warp_alive[lane] : 0u;
            u32 incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane < n) sh.warp_alive[lane] = incl - v;
            const u32 tile_alive = __shfl_sync(0xffffffffu, incl, 31);

            // 5) exclusive prefix over the previous tiles of this instance
            u32 alive_before = 0u;
#if HNB_RELAXED_ORDER
            // Reference-style order (vfx_update.wgsl:164): one aggregated atomic per tile.
            if (lane == 0) alive_before = atomicAdd(&P.draw_args[HNB_DRAW_INDEXED_INDIRECT_STRIDE * md->indirect_render_index + 1u], tile_alive);
            alive_before = __shfl_sync(0xffffffffu, alive_before, 0);
#else
            u64* states = P.tile_state;
            if (tile_in_effect == 0u) {
                if (lane == 0) hnb_st_state(&states[tile], hnb_pack_state(epoch, HNB_FLAG_PREFIX, tile_alive));
            } else {
                if (lane == 0) hnb_st_state(&states[tile], hnb_pack_state(epoch, HNB_FLAG_AGGREGATE, tile_alive));
                const u32 first_tile = tile - tile_in_effect;
                u32 pos = tile - 1u;  // newest predecessor examined by lane 0
                for (;;) {
                    const bool in_range = (pos >= first_tile + lane) && (pos >= lane);
                    u64 s = 0;
                    bool ready = true, is_prefix = true;
                    if (in_range) {
                        s = hnb_ld_state(&states[pos - lane]);
                        ready = (u32(s >> 34) == epoch) && ((s >> 32) & 3ull) != 0ull;
                        is_prefix = ((s >> 32) & 3ull) == HNB_FLAG_PREFIX;
                    }
                    // lanes past the first tile of the instance act as a PREFIX of 0 (terminates the walk)
                    const u32 ready_mask = __ballot_sync(0xffffffffu, ready);
                    const u32 prefix_mask = __ballot_sync(0xffffffffu, ready && is_prefix);
                    const u32 first_p = prefix_mask ? (u32)(__ffs(prefix_mask) - 1) : 32u;
                    const u32 need = first_p >= 31u ? 0xffffffffu : ((2u << first_p) - 1u);
                    if ((ready_mask & need) != need) continue;  // a needed predecessor has not published yet
                    u32 contrib = (in_range && lane <= first_p) ? u32(s) : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                    alive_before += contrib;
                    if (prefix_mask) break;
                    pos -= 32u;
                }
                if (lane == 0) hnb_st_state(&states[tile], hnb_pack_state(epoch, HNB_FLAG_PREFIX, alive_before + tile_alive));
            }
#endif
            if (lane == 0) {
                sh.alive_before = alive_before;
                sh.tile_alive = tile_alive;
            }
        }
        __syncthreads();
        const u32 alive_before = sh.alive_before;

        // 6) compact: survivors into the write list, the dead onto the dead stack (:148-166)
#pragma unroll
        for (int k = 0; k < HNB_TILE_K; ++k) {
            if (valid[k]) {
                const u32 row = row0 + k * HNB_BLOCK + tid;
                // number of surviving rows before `row` in this instance
                const u32 alive_rank = alive_before + sh.warp_alive[k * HNB_WARPS + warp] + rank[k];
                if (alive[k]) {
                    write_col[alive_rank] = pidx[k];
                } else {
#if HNB_RELAXED_ORDER
                    const u32 alive_index = atomicSub(&md->alive_count, 1u) - 1u;
                    P.slab.dead_index[base_particle + alive_index] = base_particle + pidx[k];
                    atomicAdd(&md->max_spawn, 1u);
#else
                    // `row - alive_rank` dead rows precede this one: serial-order value of
                    // atomicSub(alive_count,1)-1 given alive_count == max_update at pass start.
                    const u32 alive_index = max_update - 1u - (row - alive_rank);
                    P.slab.dead_index[base_particle + alive_index] = base_particle + pidx[k];
#endif
                }
            }
        }
#if !HNB_RELAXED_ORDER
        // 7) the last tile of the instance publishes the totals (replaces the per-particle atomics on
        //    instance_count / alive_count / max_spawn)
        if (tid == 0 && row0 + HNB_TILE >= max_update) {
            const u32 alive_total = alive_before + sh.tile_alive;
            const u32 dead_total = max_update - alive_total;
            P.draw_args[HNB_DRAW_INDEXED_INDIRECT_STRIDE * md->indirect_render_index + 1u] = alive_total;
            md->alive_count = md->alive_count - dead_total;
            md->max_spawn = md->max_spawn + dead_total;
        }
#endif
    }
}

}  // namespace hnb
