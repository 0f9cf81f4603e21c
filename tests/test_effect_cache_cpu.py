"""Slab slice allocator and effect cache replayed on the reference's own tests:
src/render/effect_cache.rs `effect_buffer` (:1355-1424), `pop_free_slice` (:1426-1496), `effect_cache` (:1498-1560)."""
from bevy_hanabi_b200.cache import SLAB_FREE, SLAB_MIN_CAPACITY, SLAB_USED, EffectCache, SliceAllocator


def test_effect_buffer():
    capacity = 4096
    buffer = SliceAllocator(capacity)
    assert buffer.capacity == max(capacity, SLAB_MIN_CAPACITY)
    assert buffer.used_size == 0
    assert buffer.free_slices == []
    assert buffer.allocate(buffer.capacity + 1) is None

    offset, slices = 0, []
    for size in [32, 128, 55, 148, 1, 2048, 42]:
        s = buffer.allocate(size)
        assert s == range(offset, offset + size)
        slices.append(s)
        offset += size
    assert buffer.used_size == offset

    assert buffer.free_slice(slices[2]) == SLAB_USED
    assert buffer.free_slices == [range(160, 215)]
    assert buffer.used_size == offset  # didn't move

    for k in (3, 4, 5):
        assert buffer.free_slice(slices[k]) == SLAB_USED
    assert len(buffer.free_slices) == 4
    assert buffer.used_size == offset

    # collapses all the way down to slices[1], the highest allocated
    assert buffer.free_slice(slices[6]) == SLAB_USED
    assert buffer.free_slices == []
    assert buffer.used_size == 160

    assert buffer.free_slice(slices[0]) == SLAB_USED
    assert len(buffer.free_slices) == 1
    assert buffer.used_size == 160

    # collapse all, and free the buffer
    assert buffer.free_slice(slices[1]) == SLAB_FREE
    assert buffer.free_slices == []
    assert buffer.used_size == 0


def test_pop_free_slice():
    buffer = SliceAllocator(2048)
    slice0 = buffer.allocate(32)
    assert slice0 == range(0, 32) and buffer.free_slices == []
    slice1 = buffer.allocate(1024)
    assert slice1 == range(32, 1056) and buffer.free_slices == []
    assert buffer.free_slice(slice0) == SLAB_USED
    assert buffer.free_slices == [range(0, 32)]
    # larger than slice0: cannot be recycled, appended after all existing ones
    slice2 = buffer.allocate(64)
    assert slice2 == range(1056, 1120)
    assert len(buffer.free_slices) == 1
    # a small slice that fits recycles part of slice0 (split)
    assert buffer.allocate(16) == range(0, 16)
    assert buffer.free_slices == [range(16, 32)]
    # exactly the space left: recycled completely
    assert buffer.allocate(16) == range(16, 32)
    assert buffer.free_slices == []


def test_effect_cache():
    cache = EffectCache()
    assert cache.slabs() == []
    asset, capacity = 7, SLAB_MIN_CAPACITY
    effect1 = cache.insert(asset, capacity)
    assert effect1.range == range(0, capacity) and effect1.created and effect1.slab_capacity == capacity
    assert len(cache.slabs()) == 1
    # a second instance of the same effect: the first slab is full -> its own slab
    effect2 = cache.insert(asset, capacity)
    assert effect2.range == range(0, capacity) and effect2.slab_index == 1
    assert len(cache.slabs()) == 2
    # removing the first instance frees its slab; the slot stays
    assert cache.remove(effect1) == SLAB_FREE
    assert cache.slabs() == [False, True]
    # regression #60 of the reference: the freed slot is reused
    effect3 = cache.insert(asset, capacity)
    assert effect3.range == range(0, capacity) and effect3.slab_index == 0
    assert cache.slabs() == [True, True]


def test_instances_of_one_asset_share_a_slab_and_other_assets_do_not():
    """EffectCache::insert beyond the reference's test: small instances pack into one slab of the same asset (that is
    what makes them batchable, batch.rs:153-188); another asset never shares it (is_compatible, :613-621)."""
    cache = EffectCache()
    a = [cache.insert(1, 1000) for _ in range(3)]
    assert [e.slab_index for e in a] == [0, 0, 0]
    assert [e.range for e in a] == [range(0, 1000), range(1000, 2000), range(2000, 3000)]
    assert a[0].created and not a[1].created and a[0].slab_capacity == SLAB_MIN_CAPACITY
    b = cache.insert(2, 1000)
    assert b.slab_index == 1 and b.created
    assert cache.remove(a[1]) == SLAB_USED
    assert cache.insert(1, 600).range == range(1000, 1600)   # best fit into the freed slice, split
    assert cache.insert(1, 500).range == range(3000, 3500)   # the 400 rows left do not fit: bump allocation
    assert cache.insert(1, 400).range == range(1600, 2000)
