// hnb_effect_ctx.cuh — per-thread context handed to generated init/update bodies, and the names
// generated code may reference. Included between the generated type section (Particle, Properties,
// ParentParticle) and the generated function section.
//
// The reference's WGSL reaches these through module-scope variables of vfx_init.wgsl / vfx_update.wgsl
// (`sim_params`, `seed`, `particle_index`, `particle_counter`, `transform`, `properties`,
// `properties_array_index`, `is_alive`, `parent_particle`, `spawner`); CUDA C has no per-thread module
// scope, so they are fields of Ctx re-exposed under the same names by HNB_CTX_PROLOGUE.
#pragma once

namespace hnb {

struct Ctx {
    u32 seed;
    u32 particle_index;
    u32 particle_counter;
    bool is_alive;
    const SimParams* sim;
    const Properties* props;  // this instance's Properties record (index 0 == properties_array_index)
    const Spawner* spawner;
    mat4x4f transform;
    mat4x4f inverse_transform;
#if HNB_READ_PARENT
    ParentParticle parent_particle;
    u32 parent_particle_index;
#endif
#if HNB_EMIT_EVENTS
    ChildInfo* child_infos;
    u32 base_child_index;
    u32* emit_events[HNB_MAX_EVENT_BINDINGS];
    u32 emit_events_capacity[HNB_MAX_EVENT_BINDINGS];
#if HNB_ORDERED_EVENTS
    u32 event_request[HNB_MAX_EVENT_BINDINGS];  // events this particle asked for, per channel (appended later, in row order)
#endif
#endif
};

#if HNB_READ_PARENT
#define HNB_CTX_PARENT_PROLOGUE                                        \
    const ParentParticle& parent_particle = hnb_ctx.parent_particle;   \
    const u32 parent_particle_index = hnb_ctx.parent_particle_index;   \
    (void)parent_particle; (void)parent_particle_index;
#else
#define HNB_CTX_PARENT_PROLOGUE
#endif

#define HNB_CTX_PROLOGUE                                               \
    u32& seed = hnb_ctx.seed;                                          \
    bool& is_alive = hnb_ctx.is_alive;                                 \
    const SimParams& sim_params = *hnb_ctx.sim;                        \
    const u32 particle_index = hnb_ctx.particle_index;                 \
    const u32 particle_counter = hnb_ctx.particle_counter;             \
    const Properties* properties = hnb_ctx.props;                      \
    const u32 properties_array_index = 0u;                             \
    const mat4x4f& transform = hnb_ctx.transform;                      \
    const mat4x4f& inverse_transform = hnb_ctx.inverse_transform;      \
    (void)seed; (void)is_alive; (void)sim_params; (void)particle_index; (void)particle_counter; \
    (void)properties; (void)properties_array_index; (void)transform; (void)inverse_transform; \
    HNB_CTX_PARENT_PROLOGUE

#if HNB_EMIT_EVENTS
// append_spawn_events_N (reference src/lib.rs:976-993): reserve `count` entries in child N's event
// buffer with one atomic on ChildInfo.event_count, clamp to the buffer capacity, write the parent
// particle index `count` times.
HNB_DI void hnb_append_spawn_events(Ctx& hnb_ctx, u32 binding, u32 particle_index, u32 count) {
#if HNB_ORDERED_EVENTS
    // HNB_EFFECT_ORDERED_EVENTS: only record the request; k_events_* append every row's events after the update pass in
    // the canonical (row) order, so the buffer — and which appends an overflow drops — no longer depends on scheduling.
    (void)particle_index;
    hnb_ctx.event_request[binding] += count;
    return;
#endif
    if (count == 0u) return;
    const u32 capacity = hnb_ctx.emit_events_capacity[binding];
    const u32 base = min(u32(atomicAdd(&hnb_ctx.child_infos[hnb_ctx.base_child_index + binding].event_count, i32(count))), capacity);
    const u32 capped_count = min(count, capacity - base);
    for (u32 i = 0u; i < capped_count; i += 1u) hnb_ctx.emit_events[binding][base + i] = particle_index;
}
#endif

}  // namespace hnb
