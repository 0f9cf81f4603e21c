// effect_cache.cpp — host-side placement of effect instances into particle slabs:
//   * slab slice allocator  ≙ ParticleSlab::{allocate, pop_free_slice, free_slice}  (reference
//     src/render/effect_cache.rs:484-607): bump allocation, best-fit recycling of freed slices (split when
//     larger), free list kept sorted by start, trailing free slices collapse into the bump pointer;
//   * effect cache          ≙ EffectCache::{insert, remove}  (:843-930): an instance goes into the first slab of
//     the same asset with room, else into a new slab of max(capacity, 65536) rows placed in the first empty slot.
// Pure bookkeeping: the caller creates / destroys the device storage (hnb_slab_create / hnb_slab_destroy) and
// resets recycled rows (hnb_slab_reset_rows) as the returned placements say. Pinned on the reference's own tests
// (effect_cache.rs:1355-1560) in tests/test_effect_cache_cpu.py.
#include <algorithm>
#include <cstdint>
#include <optional>
#include <vector>

#include "hanabi_b200_graph.h"

namespace {

struct Range {
    uint32_t start, end;
};

struct SliceAllocator {
    uint32_t capacity = 0;
    uint32_t used_size = 0;           // bump pointer: rows [0, used_size) are allocated or on the free list
    std::vector<Range> free_slices;   // sorted by start, never adjacent to used_size

    // best fit among the freed slices; split when the slice is larger than needed
    std::optional<Range> pop_free_slice(uint32_t size) {
        size_t best = free_slices.size();
        uint32_t best_cap = UINT32_MAX;
        for (size_t i = 0; i < free_slices.size(); ++i) {
            const uint32_t cap = free_slices[i].end - free_slices[i].start;
            if (size <= cap && cap < best_cap) {
                best = i;
                best_cap = cap;
            }
        }
        if (best == free_slices.size() || best_cap == 0) return std::nullopt;
        Range r = free_slices[best];
        if (best_cap > size) {
            free_slices[best].start = r.start + size;
            r.end = r.start + size;
        } else {
            free_slices.erase(free_slices.begin() + (ptrdiff_t)best);
        }
        return r;
    }

    std::optional<Range> allocate(uint32_t size) {
        if (size > capacity) return std::nullopt;
        if (auto r = pop_free_slice(size)) return r;
        const uint64_t new_size = uint64_t(used_size) + size;
        if (new_size > capacity) return std::nullopt;
        Range r{used_size, uint32_t(new_size)};
        used_size = uint32_t(new_size);
        return r;
    }

    // true when nothing is allocated any more (SlabState::Free)
    bool free_slice(Range r) {
        if (r.end == used_size) {
            used_size = r.start;
            while (!free_slices.empty() && free_slices.back().end == used_size) {
                used_size = free_slices.back().start;
                free_slices.pop_back();
            }
            return used_size == 0;
        }
        // keep the list sorted; a range overlapping an entry is already free (the reference warns and ignores it)
        auto it = std::lower_bound(free_slices.begin(), free_slices.end(), r, [](const Range& s, const Range& x) { return s.end <= x.start; });
        if (it != free_slices.end() && it->start < r.end) return false;
        free_slices.insert(it, r);
        return false;
    }
};

}  // namespace

struct hnb_slice_allocator {
    SliceAllocator a;
};

struct hnb_effect_cache {
    struct Slab {
        bool live = false;
        uint64_t asset_id = 0;
        SliceAllocator a;
    };
    std::vector<Slab> slabs;
};

extern "C" {

hnb_slice_allocator* hnb_slice_allocator_create(uint32_t capacity) {
    auto* s = new hnb_slice_allocator();
    s->a.capacity = std::max<uint32_t>(capacity, HNB_SLAB_MIN_CAPACITY);  // ParticleSlab::new, effect_cache.rs:262
    return s;
}
void hnb_slice_allocator_destroy(hnb_slice_allocator* s) { delete s; }
uint32_t hnb_slice_allocator_capacity(const hnb_slice_allocator* s) { return s->a.capacity; }
uint32_t hnb_slice_allocator_used_size(const hnb_slice_allocator* s) { return s->a.used_size; }
uint32_t hnb_slice_allocator_free_count(const hnb_slice_allocator* s) { return (uint32_t)s->a.free_slices.size(); }
int32_t hnb_slice_allocator_free_range(const hnb_slice_allocator* s, uint32_t index, uint32_t* start, uint32_t* end) {
    if (index >= s->a.free_slices.size()) return -1;
    *start = s->a.free_slices[index].start;
    *end = s->a.free_slices[index].end;
    return 0;
}
int32_t hnb_slice_allocator_allocate(hnb_slice_allocator* s, uint32_t size, uint32_t* start, uint32_t* end) {
    auto r = s->a.allocate(size);
    if (!r) return -1;
    *start = r->start;
    *end = r->end;
    return 0;
}
int32_t hnb_slice_allocator_free(hnb_slice_allocator* s, uint32_t start, uint32_t end) { return s->a.free_slice({start, end}) ? HNB_SLAB_FREE : HNB_SLAB_USED; }

hnb_effect_cache* hnb_effect_cache_create(void) { return new hnb_effect_cache(); }
void hnb_effect_cache_destroy(hnb_effect_cache* c) { delete c; }
uint32_t hnb_effect_cache_slab_count(const hnb_effect_cache* c) { return (uint32_t)c->slabs.size(); }
int32_t hnb_effect_cache_slab_is_live(const hnb_effect_cache* c, uint32_t slab_index) { return slab_index < c->slabs.size() && c->slabs[slab_index].live ? 1 : 0; }

int32_t hnb_effect_cache_insert(hnb_effect_cache* c, uint64_t asset_id, uint32_t capacity, hnb_cached_effect* out) {
    if (!c || !out) return -1;
    // first slab of the same asset with room (is_compatible compares the asset handle only, effect_cache.rs:613-621)
    for (size_t i = 0; i < c->slabs.size(); ++i) {
        auto& s = c->slabs[i];
        if (!s.live || s.asset_id != asset_id) continue;
        if (auto r = s.a.allocate(capacity)) {
            *out = {uint32_t(i), r->start, r->end, s.a.capacity, 0u};
            return 0;
        }
    }
    // a new slab, in the first empty slot
    size_t index = c->slabs.size();
    for (size_t i = 0; i < c->slabs.size(); ++i)
        if (!c->slabs[i].live) { index = i; break; }
    if (index == c->slabs.size()) c->slabs.emplace_back();
    auto& s = c->slabs[index];
    s = hnb_effect_cache::Slab();
    s.live = true;
    s.asset_id = asset_id;
    s.a.capacity = std::max<uint32_t>(capacity, HNB_SLAB_MIN_CAPACITY);
    auto r = s.a.allocate(capacity);
    *out = {uint32_t(index), r->start, r->end, s.a.capacity, 1u};
    return 0;
}

int32_t hnb_effect_cache_remove(hnb_effect_cache* c, const hnb_cached_effect* e) {
    if (!c || !e || e->slab_index >= c->slabs.size() || !c->slabs[e->slab_index].live) return -1;
    auto& s = c->slabs[e->slab_index];
    if (s.a.free_slice({e->range_start, e->range_end})) {
        s.live = false;
        return HNB_SLAB_FREE;
    }
    return HNB_SLAB_USED;
}

}  // extern "C"
