// spawn_batch.cpp — the CPU producers that feed the hot path:
//   * SpawnerSettings / EffectSpawner::tick  (reference src/spawn.rs:219-470, :700-921): how many particles an
//     instance asks to spawn this frame -> GpuSpawnerParams.spawn;
//   * Batcher::push / EffectBatch::try_merge  (reference src/render/batch.rs:153-188, :265-386): which instances
//     share a launch, GpuBatchInfo rows and the CPU prefix sums of spawn counts consumed by the init pass.
// Pure host code, exposed through include/hanabi_b200_graph.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "hanabi_b200_graph.h"

extern "C" void hnb_set_last_error_(const char* msg);

namespace {

// Pcg32 as used by the reference's spawner RNG (rand_pcg 0.10 `Pcg32` = Lcg64Xsh32, un-vendored dependency):
// state = state * 6364136223846793005 + inc; output = rotr32(((state >> 18) ^ state) >> 27, state >> 59).
// "parity unpinned": the reference seeds it from OS entropy (spawn.rs:17-27) and no test pins sampled values;
// only CpuValue::Single paths are compared with the reference's test vectors.
struct Pcg32 {
    uint64_t state, inc;
    explicit Pcg32(uint64_t seed, uint64_t stream = 0xa02bdbf7bb3c0a7ull) {
        state = 0;
        inc = (stream << 1) | 1u;
        next();
        state += seed;
        next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    float next_f01() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
};

float sample(float lo, float hi, Pcg32& rng) {  // CpuValue::sample (spawn.rs:105-111)
    if (lo == hi) return lo;
    return lo + (hi - lo) * rng.next_f01();
}

}  // namespace

struct hnb_effect_spawner {
    hnb_spawner_settings settings;
    Pcg32 rng;
    float cycle_time = 0.f;
    uint32_t completed_cycle_count = 0;
    float sampled_spawn_duration = 0.f, sampled_period = 0.f, sampled_count = 0.f;
    uint32_t spawn_count = 0;
    float spawn_remainder = 0.f;
    bool active = true;
    hnb_effect_spawner(const hnb_spawner_settings& s, uint64_t seed) : settings(s), rng(seed) {
        // EffectSpawner::new (spawn.rs:700-718)
        const bool forever = s.cycle_count == 0;
        completed_cycle_count = (s.emit_on_start || forever) ? 0u : s.cycle_count;
        active = s.starts_active != 0;
    }
    bool is_once() const { return settings.cycle_count == 1; }
    bool is_forever() const { return settings.cycle_count == 0; }

    // EffectSpawner::tick (spawn.rs:838-921), statement by statement
    uint32_t tick(float dt) {
        if (!active || (!is_forever() && completed_cycle_count >= settings.cycle_count)) {
            spawn_count = 0;
            return 0;
        }
        for (;;) {
            if (sampled_period == 0.0f) {
                if (is_once()) {
                    sampled_spawn_duration = sample(settings.spawn_duration_lo, settings.spawn_duration_hi, rng);
                    sampled_period = std::max(sampled_spawn_duration, 1e-12f);
                } else {
                    sampled_period = sample(settings.period_lo, settings.period_hi, rng);
                    if (!(sampled_period > 0.f)) throw std::runtime_error("sampled spawner period must be > 0");
                    float d = sample(settings.spawn_duration_lo, settings.spawn_duration_hi, rng);
                    sampled_spawn_duration = std::min(std::max(d, 0.f), sampled_period);
                }
                // the reference samples spawn_duration a second time and that sample wins (spawn.rs:861-867)
                sampled_spawn_duration = sample(settings.spawn_duration_lo, settings.spawn_duration_hi, rng);
                sampled_count = std::max(sample(settings.count_lo, settings.count_hi, rng), 0.f);
            }
            const float new_time = cycle_time + dt;
            if (cycle_time <= sampled_spawn_duration) {
                if (sampled_spawn_duration < std::max(1e-5f, dt / 100.0f)) {
                    spawn_remainder += sampled_count;
                } else {
                    float ratio = (std::min(new_time, sampled_spawn_duration) - cycle_time) / sampled_spawn_duration;
                    ratio = std::min(std::max(ratio, 0.f), 1.f);
                    spawn_remainder += sampled_count * ratio;
                }
            }
            cycle_time = new_time;
            if (cycle_time >= sampled_period) {
                dt = cycle_time - sampled_period;
                cycle_time = 0.0f;
                completed_cycle_count += 1;
                sampled_period = 0.0f;
                if (!is_forever() && completed_cycle_count >= settings.cycle_count) break;
            } else {
                break;
            }
        }
        const float count = std::floor(spawn_remainder);
        spawn_remainder -= count;
        spawn_count = (uint32_t)count;
        return spawn_count;
    }
    void reset() {
        cycle_time = 0.f;
        completed_cycle_count = 0;
        sampled_spawn_duration = sampled_period = sampled_count = 0.f;
        spawn_count = 0;
        spawn_remainder = 0.f;
    }
};

// ---- Batcher (batch.rs) -------------------------------------------------------------------------
struct hnb_batcher {
    struct Batch {
        hnb_batch_key key;
        uint32_t total_spawn_count;
        uint32_t num_instances;
    };
    std::vector<Batch> batches;
    std::vector<hnb_batch_info> infos;
    std::vector<uint32_t> prefix;
    uint32_t cpu_prefix_sum_value = 0;
    bool open = false;

    static bool can_merge(const hnb_batch_key& a, const hnb_batch_key& b) {
        // EffectBatch::try_merge (batch.rs:153-173): conservative — same asset, slab, pipelines, property buffer,
        // parent; no GPU-event effect is ever merged; both must be CPU-spawned
        return a.asset_id == b.asset_id && a.slab_id == b.slab_id && a.pipeline_id == b.pipeline_id && a.property_key == b.property_key &&
               a.parent_slab_id == b.parent_slab_id && !a.uses_gpu_events && !b.uses_gpu_events && a.is_cpu_spawner && b.is_cpu_spawner;
    }
    void end_batch() {
        if (!open) return;
        infos.back().prefix_sum_count = (uint32_t)prefix.size() - infos.back().prefix_sum_offset;
        open = false;
    }
    // Batcher::push (batch.rs:348-386). Returns the new batch index, or -1 when merged into the previous batch.
    int32_t push(const hnb_batch_key& key, uint32_t spawner_base, uint32_t slab_offset, uint32_t instance_spawn_count) {
        if (!batches.empty() && can_merge(batches.back().key, key)) {
            prefix.push_back(cpu_prefix_sum_value);
            cpu_prefix_sum_value += instance_spawn_count;
            batches.back().total_spawn_count += key.is_cpu_spawner ? instance_spawn_count : 0;
            batches.back().num_instances++;
            return -1;
        }
        end_batch();
        batches.push_back({key, key.is_cpu_spawner ? instance_spawn_count : 0u, 1u});
        hnb_batch_info bi{};
        bi.total_spawn_count = 0;  // written as 0 and never updated in the reference (batch.rs:271-278)
        bi.total_update_count = 0;
        bi.spawner_base = spawner_base;
        bi.base_particle = slab_offset;
        bi.prefix_sum_offset = (uint32_t)prefix.size();
        bi.prefix_sum_count = 0xFFFFFFFFu;
        infos.push_back(bi);
        open = true;
        cpu_prefix_sum_value = 0;
        prefix.push_back(cpu_prefix_sum_value);
        cpu_prefix_sum_value += instance_spawn_count;
        return (int32_t)batches.size() - 1;
    }
};

namespace {
template <typename F> int32_t guarded(F&& f) {
    try {
        f();
        return HNB_OK;
    } catch (const std::exception& e) {
        hnb_set_last_error_(e.what());
        return HNB_ERR_INVALID_ARG;
    }
}
}  // namespace

extern "C" {

int32_t hnb_spawner_settings_new(float count_lo, float count_hi, float duration_lo, float duration_hi, float period_lo, float period_hi,
                                 uint32_t cycle_count, hnb_spawner_settings* out) {
    return guarded([&] {
        if (!out) throw std::invalid_argument("out is NULL");
        // SpawnerSettings::try_new (spawn.rs:313-338)
        const float pmin = std::min(period_lo, period_hi), pmax = std::max(period_lo, period_hi);
        if (cycle_count != 1 && (pmin < 0.f || pmax <= 0.f)) {
            if (pmin < 0.f) throw std::invalid_argument("`period` must not generate negative numbers (period.min was " + std::to_string(pmin) + ", expected >= 0).");
            throw std::invalid_argument("`period` must be able to generate a positive number (period.max was " + std::to_string(pmax) + ", expected > 0).");
        }
        if (!std::isfinite(pmin) || !std::isfinite(pmax)) throw std::invalid_argument("`period` has an infinite bound; use cycle_count = 1 for a single-cycle burst.");
        *out = hnb_spawner_settings{count_lo, count_hi, duration_lo, duration_hi, period_lo, period_hi, cycle_count, 1u, 1u};
    });
}
int32_t hnb_spawner_settings_once(float count, hnb_spawner_settings* out) { return hnb_spawner_settings_new(count, count, 0.f, 0.f, 0.f, 0.f, 1, out); }
int32_t hnb_spawner_settings_rate(float rate, hnb_spawner_settings* out) { return hnb_spawner_settings_new(rate, rate, 1.f, 1.f, 1.f, 1.f, 0, out); }
int32_t hnb_spawner_settings_burst(float count, float period, hnb_spawner_settings* out) {
    return hnb_spawner_settings_new(count, count, 0.f, 0.f, period, period, 0, out);
}

hnb_effect_spawner* hnb_effect_spawner_create(const hnb_spawner_settings* settings, uint64_t rng_seed) {
    if (!settings) return nullptr;
    return new hnb_effect_spawner(*settings, rng_seed);
}
void hnb_effect_spawner_destroy(hnb_effect_spawner* s) { delete s; }
int32_t hnb_effect_spawner_tick(hnb_effect_spawner* s, float dt, uint32_t* spawn_count) {
    return guarded([&] {
        uint32_t n = s->tick(dt);
        if (spawn_count) *spawn_count = n;
    });
}
void hnb_effect_spawner_reset(hnb_effect_spawner* s) { s->reset(); }
void hnb_effect_spawner_set_active(hnb_effect_spawner* s, uint32_t active) { s->active = active != 0; }
int32_t hnb_effect_spawner_state(const hnb_effect_spawner* s, hnb_effect_spawner_state_t* out) {
    return guarded([&] {
        if (!out) throw std::invalid_argument("out is NULL");
        out->cycle_time = s->cycle_time;
        out->cycle_spawn_duration = s->sampled_spawn_duration;
        out->cycle_period = s->is_once() ? 0.f : s->sampled_period;
        out->cycle_ratio = s->is_once() ? 0.f : s->cycle_time / s->sampled_period;
        out->cycle_spawn_count = s->sampled_count;
        out->completed_cycle_count = s->completed_cycle_count;
        out->active = s->active ? 1u : 0u;
        out->has_completed = (!s->is_forever() && s->completed_cycle_count >= s->settings.cycle_count) ? 1u : 0u;
        out->spawn_count = s->spawn_count;
    });
}

hnb_batcher* hnb_batcher_create(void) { return new hnb_batcher(); }
void hnb_batcher_destroy(hnb_batcher* b) { delete b; }
void hnb_batcher_clear(hnb_batcher* b) {
    b->batches.clear();
    b->infos.clear();
    b->prefix.clear();
    b->cpu_prefix_sum_value = 0;
    b->open = false;
}
int32_t hnb_batcher_push(hnb_batcher* b, const hnb_batch_key* key, uint32_t spawner_base, uint32_t slab_offset, uint32_t instance_spawn_count,
                         int32_t* batch_index) {
    return guarded([&] {
        if (!key) throw std::invalid_argument("key is NULL");
        int32_t idx = b->push(*key, spawner_base, slab_offset, instance_spawn_count);
        if (batch_index) *batch_index = idx;
    });
}
int32_t hnb_batcher_finish(hnb_batcher* b, const hnb_batch_info** infos, uint32_t* n_batches, const uint32_t** prefix, uint32_t* n_prefix,
                           uint32_t* total_spawn_counts, uint32_t total_cap) {
    return guarded([&] {
        b->end_batch();
        if (infos) *infos = b->infos.data();
        if (n_batches) *n_batches = (uint32_t)b->infos.size();
        if (prefix) *prefix = b->prefix.data();
        if (n_prefix) *n_prefix = (uint32_t)b->prefix.size();
        for (size_t i = 0; i < b->batches.size() && i < total_cap && total_spawn_counts; ++i) total_spawn_counts[i] = b->batches[i].total_spawn_count;
    });
}

}  // extern "C"

// ---- EffectSorter (batch.rs:476-637) ------------------------------------------------------------
// Order in which batch_effects() walks the instances: dependency level first (children get level 0, an effect
// sits one level above its deepest child — so children come BEFORE their parents, as the reference's own test
// pins, batch.rs:776-826), then slab, then position in the slab, which puts mergeable instances next to each other.
struct hnb_effect_sorter {
    struct Entry {
        uint64_t entity;
        uint32_t slab_id, base_instance;
    };
    std::vector<Entry> effects;
    std::vector<std::pair<uint64_t, uint64_t>> child_to_parent;  // (child, parent); a child has at most one parent

    bool sort() {
        const size_t n = effects.size();
        auto index_of = [&](uint64_t e) -> size_t {
            for (size_t i = 0; i < n; ++i)
                if (effects[i].entity == e) return i;
            return n;
        };
        std::vector<std::vector<size_t>> parents(n);
        for (auto& cp : child_to_parent) {
            const size_t kid = index_of(cp.first), parent = index_of(cp.second);
            if (kid == n || parent == n) return false;
            parents[kid].push_back(parent);
        }
        // depth-first topological order: an effect is emitted after the parents it points to
        std::vector<size_t> ordering;
        std::vector<uint8_t> state(n, 0);  // 0 new, 1 visiting, 2 done
        bool cycle = false;
        struct Frame { size_t node, next; };
        for (size_t root = 0; root < n; ++root) {
            if (state[root]) continue;
            std::vector<Frame> stack{{root, 0}};
            state[root] = 1;
            while (!stack.empty()) {
                Frame& f = stack.back();
                if (f.next < parents[f.node].size()) {
                    const size_t p = parents[f.node][f.next++];
                    if (state[p] == 1) cycle = true;
                    if (state[p] == 0) { state[p] = 1; stack.push_back({p, 0}); }
                } else {
                    state[f.node] = 2;
                    ordering.push_back(f.node);
                    stack.pop_back();
                }
            }
        }
        if (cycle) return false;
        std::vector<uint32_t> levels(n, 0);
        for (size_t k = ordering.size(); k-- > 0;) {
            const size_t e = ordering[k];
            for (size_t p : parents[e]) levels[p] = std::max(levels[p], levels[e] + 1);
        }
        std::vector<size_t> perm(n);
        for (size_t i = 0; i < n; ++i) perm[i] = i;
        std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
            if (levels[a] != levels[b]) return levels[a] < levels[b];
            if (effects[a].slab_id != effects[b].slab_id) return effects[a].slab_id < effects[b].slab_id;
            return effects[a].base_instance < effects[b].base_instance;
        });
        std::vector<Entry> sorted(n);
        for (size_t i = 0; i < n; ++i) sorted[i] = effects[perm[i]];
        effects.swap(sorted);
        return true;
    }
};

extern "C" {
hnb_effect_sorter* hnb_effect_sorter_create(void) { return new hnb_effect_sorter(); }
void hnb_effect_sorter_destroy(hnb_effect_sorter* s) { delete s; }
void hnb_effect_sorter_insert(hnb_effect_sorter* s, uint64_t entity, uint32_t slab_id, uint32_t base_instance, uint64_t parent) {
    s->effects.push_back({entity, slab_id, base_instance});
    if (parent != HNB_NO_ENTITY) {
        for (auto& cp : s->child_to_parent)
            if (cp.first == entity) { cp.second = parent; return; }
        s->child_to_parent.push_back({entity, parent});
    }
}
int32_t hnb_effect_sorter_sort(hnb_effect_sorter* s) {
    if (s->sort()) return 0;
    hnb_set_last_error_("EffectSorter: unknown parent entity or cyclic parent-child relation");
    return -1;
}
uint32_t hnb_effect_sorter_len(const hnb_effect_sorter* s) { return (uint32_t)s->effects.size(); }
uint64_t hnb_effect_sorter_get(const hnb_effect_sorter* s, uint32_t index) { return index < s->effects.size() ? s->effects[index].entity : HNB_NO_ENTITY; }
}
