"""Host-side placement of effect instances into slabs: the slice allocator of one slab
(≙ ParticleSlab::allocate / free_slice, reference src/render/effect_cache.rs:484-607) and the effect cache
(≙ EffectCache::insert / remove, :843-938). Bookkeeping only; device storage is created by the caller
(`Context.slab_create`) from what `EffectCache.insert` returns."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from ._native import lib

u32, P = C.c_uint32, C.POINTER
SLAB_MIN_CAPACITY = 65536
SLAB_USED, SLAB_FREE = 0, 1


class CachedEffectC(C.Structure):
    _fields_ = [("slab_index", u32), ("range_start", u32), ("range_end", u32), ("slab_capacity", u32), ("created", u32)]


CACHE_SIGNATURES = {
    "hnb_slice_allocator_create": (C.c_void_p, [u32]),
    "hnb_slice_allocator_destroy": (None, [C.c_void_p]),
    "hnb_slice_allocator_capacity": (u32, [C.c_void_p]),
    "hnb_slice_allocator_used_size": (u32, [C.c_void_p]),
    "hnb_slice_allocator_free_count": (u32, [C.c_void_p]),
    "hnb_slice_allocator_free_range": (C.c_int32, [C.c_void_p, u32, P(u32), P(u32)]),
    "hnb_slice_allocator_allocate": (C.c_int32, [C.c_void_p, u32, P(u32), P(u32)]),
    "hnb_slice_allocator_free": (C.c_int32, [C.c_void_p, u32, u32]),
    "hnb_effect_cache_create": (C.c_void_p, []),
    "hnb_effect_cache_destroy": (None, [C.c_void_p]),
    "hnb_effect_cache_slab_count": (u32, [C.c_void_p]),
    "hnb_effect_cache_slab_is_live": (C.c_int32, [C.c_void_p, u32]),
    "hnb_effect_cache_insert": (C.c_int32, [C.c_void_p, C.c_uint64, u32, P(CachedEffectC)]),
    "hnb_effect_cache_remove": (C.c_int32, [C.c_void_p, P(CachedEffectC)]),
}
for _n, (_r, _a) in CACHE_SIGNATURES.items():
    _f = getattr(lib, _n)
    _f.restype, _f.argtypes = _r, _a


class SliceAllocator:
    def __init__(self, capacity: int):
        self._h = lib.hnb_slice_allocator_create(capacity)

    def __del__(self):
        try:
            if self._h:
                lib.hnb_slice_allocator_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def capacity(self) -> int:
        return lib.hnb_slice_allocator_capacity(self._h)

    @property
    def used_size(self) -> int:
        return lib.hnb_slice_allocator_used_size(self._h)

    @property
    def free_slices(self) -> list[range]:
        out = []
        for i in range(lib.hnb_slice_allocator_free_count(self._h)):
            a, b = u32(), u32()
            lib.hnb_slice_allocator_free_range(self._h, i, C.byref(a), C.byref(b))
            out.append(range(a.value, b.value))
        return out

    def allocate(self, size: int) -> range | None:
        a, b = u32(), u32()
        if lib.hnb_slice_allocator_allocate(self._h, size, C.byref(a), C.byref(b)) != 0:
            return None
        return range(a.value, b.value)

    def free_slice(self, r: range) -> int:
        return lib.hnb_slice_allocator_free(self._h, r.start, r.stop)


@dataclass
class CachedEffect:
    slab_index: int
    range: range
    slab_capacity: int
    created: bool


class EffectCache:
    def __init__(self):
        self._h = lib.hnb_effect_cache_create()

    def __del__(self):
        try:
            if self._h:
                lib.hnb_effect_cache_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def slabs(self) -> list[bool]:
        """One entry per slot, True when the slot holds a live slab (≙ `slabs()[i].is_some()`)."""
        return [bool(lib.hnb_effect_cache_slab_is_live(self._h, i)) for i in range(lib.hnb_effect_cache_slab_count(self._h))]

    def insert(self, asset_id: int, capacity: int) -> CachedEffect:
        c = CachedEffectC()
        if lib.hnb_effect_cache_insert(self._h, asset_id, capacity, C.byref(c)) != 0:
            raise RuntimeError("hnb_effect_cache_insert failed")
        return CachedEffect(c.slab_index, range(c.range_start, c.range_end), c.slab_capacity, bool(c.created))

    def remove(self, e: CachedEffect) -> int:
        c = CachedEffectC(e.slab_index, e.range.start, e.range.stop, e.slab_capacity, int(e.created))
        state = lib.hnb_effect_cache_remove(self._h, C.byref(c))
        if state < 0:
            raise KeyError("no such cached effect")
        return state
