"""The effect-independent kernels under the CPU emulation (tests/static_emu.py): bookkeeping (fused and un-fused) against
the C oracle's restatement of vfx_indirect / vfx_prefix_sum, and the two ribbon-sort kernels against a stable sort."""
import ctypes as C

import numpy as np
import pytest

from oracle import c_oracle as O
from tests import static_emu as S

u32 = np.uint32
u32p = C.POINTER(C.c_uint32)
pytestmark = pytest.mark.timeout(600)  # real threads: a protocol bug must fail, not hang


@pytest.fixture(scope="module")
def semu():
    return S.build()


def _tables(rng, batches, tile_sizes):
    """batches: list of instance counts. Random alive / capacity / spawn state per instance."""
    n = sum(batches)
    md = (O.EffectMetadata * n)()
    sp = (O.Spawner * n)()
    draw = rng.integers(0, 100, 5 * n).astype(u32)
    spawn_range = np.zeros(n, dtype=u32)
    for i in range(n):
        cap = int(rng.integers(1, 5000))
        alive = int(rng.integers(0, cap + 1))
        md[i].capacity, md[i].alive_count, md[i].max_spawn = cap, alive, cap - alive
        md[i].indirect_write_index, md[i].indirect_render_index, md[i].particle_counter = int(rng.integers(0, 2)), i, int(rng.integers(0, 1000))
        spawn = int(rng.integers(-3, 60)) if rng.random() < 0.6 else 0
        sp[i].spawn, sp[i].effect_metadata_index, sp[i].draw_indirect_index = spawn, i, i
        spawn_range[i] = max(spawn, 0) + (int(rng.integers(0, 64)) if spawn > 0 else 0)   # threads mapped >= spawn count
    bis = (O.BatchInfo * len(batches))()
    off = 0
    for b, cnt in enumerate(batches):
        bis[b].spawner_base = bis[b].prefix_sum_offset = off
        bis[b].prefix_sum_count = cnt
        off += cnt
    frame = np.zeros(16, dtype=u32)
    frame[6], frame[7], frame[8] = n, 1, len(batches)   # sim.num_effects | epoch | num_batches
    return dict(n=n, md=md, sp=sp, draw=draw, spawn_range=spawn_range, bis=bis, frame=frame, prefix=np.zeros(n, dtype=u32),
                tile_prefix=np.zeros(n + 1, dtype=u32), tile_size=np.array(tile_sizes, dtype=u32), dispatch=np.zeros(3 * len(batches), dtype=u32),
                batch_tiles=np.zeros(len(batches), dtype=u32), tickets=np.full(len(batches), 77, dtype=u32))


def _bind(t):
    T = S.StaticTables()
    p = lambda a: a.ctypes.data
    T.frame, T.spawners, T.spawn_range, T.prefix_sum, T.tile_prefix = p(t["frame"]), C.addressof(t["sp"]), p(t["spawn_range"]), p(t["prefix"]), p(t["tile_prefix"])
    T.batch_infos, T.batch_tile_size, T.dispatch_args, T.batch_tiles, T.tickets = C.addressof(t["bis"]), p(t["tile_size"]), p(t["dispatch"]), p(t["batch_tiles"]), p(t["tickets"])
    T.metadata, T.draw_args, T.child_infos, T.num_child_infos = C.addressof(t["md"]), p(t["draw"]), None, 0
    return T


def _expected(orc, t):
    """Deferred init accounting restated + the C oracle's vfx_indirect / vfx_prefix_sum + the tile prefix in numpy."""
    n, md, sp = t["n"], (O.EffectMetadata * t["n"]).from_buffer_copy(bytes(t["md"])), (O.Spawner * t["n"]).from_buffer_copy(bytes(t["sp"]))
    for i in range(n):
        if t["spawn_range"][i]:
            passed = min(int(t["spawn_range"][i]), max(sp[i].spawn, 0) if sp[i].spawn >= 0 else int(np.uint32(sp[i].spawn)), md[i].max_spawn)
            md[i].alive_count += passed
            md[i].particle_counter += passed
    draw, prefix = t["draw"].copy(), np.zeros(n, dtype=u32)
    sim = O.SimParams(0, 0, 0, 0, 0, 0, n)
    orc.orc_indirect(C.byref(sim), md, draw.ctypes.data_as(u32p), sp, prefix.ctypes.data_as(u32p), None, 0)
    alive = prefix.copy()
    bis = (O.BatchInfo * len(t["bis"])).from_buffer_copy(bytes(t["bis"]))
    dispatch = np.zeros(3 * len(bis), dtype=u32)
    orc.orc_prefix_sum(bis, len(bis), prefix.ctypes.data_as(u32p), dispatch.ctypes.data_as(u32p))
    tile_prefix, batch_tiles = np.zeros(n, dtype=u32), np.zeros(len(bis), dtype=u32)
    for b in range(len(bis)):
        lo, cnt = bis[b].prefix_sum_offset, bis[b].prefix_sum_count
        tiles = (alive[lo:lo + cnt] + t["tile_size"][b] - 1) // t["tile_size"][b]
        tile_prefix[lo:lo + cnt] = np.concatenate([[0], np.cumsum(tiles)[:-1]]) if cnt else []
        batch_tiles[b] = tiles.sum()
    return dict(md=bytes(md), sp=bytes(sp), draw=draw, prefix=prefix, bis=bytes(bis), dispatch=dispatch, tile_prefix=tile_prefix, batch_tiles=batch_tiles)


def _check(t, want):
    n = t["n"]
    assert bytes(t["md"]) == want["md"]
    assert bytes(t["sp"]) == want["sp"]
    np.testing.assert_array_equal(t["draw"], want["draw"])
    np.testing.assert_array_equal(t["prefix"], want["prefix"])
    assert bytes(t["bis"]) == want["bis"]
    np.testing.assert_array_equal(t["dispatch"], want["dispatch"])
    np.testing.assert_array_equal(t["tile_prefix"][:n], want["tile_prefix"])
    np.testing.assert_array_equal(t["batch_tiles"], want["batch_tiles"])
    assert not t["tickets"].any() and not t["spawn_range"].any()


@pytest.mark.parametrize("batches", [[1], [3, 1, 40], [700], [300, 2, 257, 256]])
def test_fused_bookkeeping_kernel(semu, orc, batches):
    """k_bookkeeping (one CTA per batch, block scans over chunks of 256 instances, carries) == indirect + prefix sum."""
    rng = np.random.default_rng(sum(batches))
    t = _tables(rng, batches, [128 * (1 + b % 3) for b in range(len(batches))])
    want = _expected(orc, t)
    semu.semu_bookkeeping(C.byref(_bind(t)), len(batches))
    _check(t, want)


def test_unfused_passes_equal_fused(semu, orc):
    rng = np.random.default_rng(9)
    t = _tables(rng, [5, 130, 1], [128, 256, 512])
    want = _expected(orc, t)
    T = _bind(t)
    semu.semu_indirect(C.byref(T), t["n"])
    semu.semu_prefix_sum(C.byref(T), 3)
    _check(t, want)
    # the stand-alone tile-prefix kernel rebuilds one batch's table for another tile size from max_update
    semu.semu_tile_prefix(C.byref(T), 1, 64)
    md = (O.EffectMetadata * t["n"]).from_buffer_copy(want["md"])
    tiles = np.array([(md[i].max_update + 63) // 64 for i in range(5, 135)], dtype=u32)
    np.testing.assert_array_equal(t["tile_prefix"][5:135], np.concatenate([[0], np.cumsum(tiles)[:-1]]))
    assert t["batch_tiles"][1] == tiles.sum()


# ---- ribbon sort ------------------------------------------------------------------------------------
def _sort_world(rng, counts, wide):
    """One slab, instances back to back with capacity = count + 9; particle records of 12 words in three planes
    (16 + 16 + 16 bytes), ribbon id in word 9, age in word 3."""
    caps = [c + 9 for c in counts]
    rows, n = sum(caps), len(counts)
    planes = [rng.integers(0, 2**32, rows * 4).astype(u32) for _ in range(3)]
    K1, K2 = 9, 3
    if not wide:
        planes[2][1::4] = rng.integers(0, 4, rows)                                            # word 9 = plane 2, lane 1
        planes[0][3::4] = rng.choice(np.array([0.0, 0.5, 1.0, 2.5], dtype=np.float32), rows).view(u32)   # word 3 = plane 0, lane 3
    ping, pong = rng.integers(0, 2**32, rows).astype(u32), rng.integers(0, 2**32, rows).astype(u32)
    md, sp = (O.EffectMetadata * n)(), (O.Spawner * n)()
    off, expect = 0, []
    for i, (c, cap) in enumerate(zip(counts, caps)):
        col = int(rng.integers(0, 2))
        perm = rng.permutation(cap)[:c].astype(u32)
        (ping if col == 0 else pong)[off:off + c] = perm
        md[i].capacity, md[i].alive_count, md[i].indirect_write_index, md[i].sort_key_offset, md[i].sort_key2_offset, md[i].particle_stride = cap, c, col, K1, K2, 12
        sp[i].effect_metadata_index, sp[i].slab_offset = i, off
        r = off + perm.astype(np.int64)
        order = np.lexsort((planes[0][r * 4 + 3], planes[2][r * 4 + 1]))
        expect.append((off, c, col, perm[order]))
        off += cap
    a = S.RibbonSortArgs()
    for p in range(3):
        a.planes.ptr[p], a.planes.words[p], a.planes.word_off[p] = planes[p].ctypes.data, 4, 4 * p
        for w in range(4):
            a.planes.word_to_plane[4 * p + w] = p
    a.ping, a.pong, a.spawners, a.metadata = ping.ctypes.data, pong.ctypes.data, C.addressof(sp), C.addressof(md)
    a.spawner_base, a.instance_count = 0, n
    keep = (planes, ping, pong, md, sp)
    return a, keep, expect, rows


def _check_sorted(keep, expect, before):
    _, ping, pong, _, _ = keep
    want_ping, want_pong = before
    for off, c, col, sorted_vals in expect:
        (want_ping if col == 0 else want_pong)[off:off + c] = sorted_vals
    np.testing.assert_array_equal(ping, want_ping)
    np.testing.assert_array_equal(pong, want_pong)


@pytest.mark.parametrize("wide", [False, True])
def test_ribbon_sort_small_kernel(semu, wide):
    rng = np.random.default_rng(int(wide))
    a, keep, expect, _ = _sort_world(rng, [0, 1, 2, 3, 31, 100, 777, 2047, 2048], wide)
    before = (keep[1].copy(), keep[2].copy())
    semu.semu_ribbon_sort_small(C.byref(a))
    _check_sorted(keep, expect, before)


@pytest.mark.parametrize("wide", [False, True])
def test_ribbon_sort_large_kernel(semu, wide):
    """Cooperative radix sort on a grid of 3 CTAs: chunking, digit-major offsets, skipped passes (narrow keys),
    all eight passes (wide keys), two large instances in one launch (histogram double-buffering), small ones skipped."""
    rng = np.random.default_rng(10 + int(wide))
    a, keep, expect, rows = _sort_world(rng, [2049, 50, 7000, 0, 3333], wide)
    grid = 3
    keys = [np.zeros(rows, dtype=np.uint64) for _ in range(2)]
    vals = [np.zeros(rows, dtype=u32) for _ in range(2)]
    hist = np.zeros(semu.semu_hist_words(grid), dtype=u32)
    for i in range(2):
        a.scratch_keys[i], a.scratch_vals[i] = keys[i].ctypes.data, vals[i].ctypes.data
    a.scratch_hist, a.scratch_rows = hist.ctypes.data, rows
    before = (keep[1].copy(), keep[2].copy())
    semu.semu_ribbon_sort_large(C.byref(a), grid)
    expect_large = [e for e in expect if e[1] > 2048]     # the large kernel leaves n <= 2048 to the small one
    _check_sorted(keep, expect_large, before)


# ---- ordered event append ---------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cap", [(0, 16), (1, 16), (300, 64), (2048, 256), (2049, 100000), (9000, 1000)])
def test_ordered_event_append_kernels(semu, rows, cap):
    """k_events_block_sums / _scan_blocks / _write against the serial definition (append_spawn_events_N in thread order,
    lib.rs:976-993): position = exclusive prefix of the requested counts, clamped to the buffer; count = total."""
    rng = np.random.default_rng(rows + cap)
    capacity_rows, base = rows + 500, 37
    counts = rng.integers(0, 2**32, capacity_rows).astype(u32)                  # junk beyond `rows` must be ignored
    counts[:rows] = np.where(rng.random(rows) < 0.3, rng.integers(1, 6, rows), 0)
    if rows > 10:
        counts[rows // 2] = 700                                                  # one particle asking for many events
    read_col = rng.permutation(capacity_rows).astype(u32)
    ping, pong = np.zeros(base + capacity_rows, dtype=u32), np.zeros(base + capacity_rows, dtype=u32)
    pong[base:] = read_col                                                       # indirect_write_index 0 -> the update read pong
    md, sp = (O.EffectMetadata * 1)(), (O.Spawner * 1)()
    md[0].max_update, md[0].indirect_write_index, md[0].base_child_index = rows, 0, 2
    sp[0].slab_offset = base
    child_infos = np.zeros((5, 2), dtype=np.int32)
    child_infos[3, 1] = 11                                                       # event_count accumulates (atomicAdd)
    buffer = np.full(cap, 0xDEADBEEF, dtype=u32)
    block_sums = np.zeros(capacity_rows // 2048 + 2, dtype=u32)
    a = S.EventAppendArgs()
    a.counts, a.ping, a.pong, a.spawner, a.metadata = counts.ctypes.data, ping.ctypes.data, pong.ctypes.data, C.addressof(sp), C.addressof(md)
    a.block_sums, a.child_infos, a.binding, a.buffer, a.capacity = block_sums.ctypes.data, child_infos.ctypes.data, 1, buffer.ctypes.data, cap
    semu.semu_ordered_event_append(C.byref(a), capacity_rows)
    want = np.repeat(read_col[:rows], counts[:rows])
    total = int(counts[:rows].sum())
    assert child_infos[3, 1] == 11 + total and not child_infos[[0, 1, 2, 4]].any()
    kept = min(total, cap)
    np.testing.assert_array_equal(buffer[:kept], want[:kept])
    assert (buffer[kept:] == 0xDEADBEEF).all()


def test_ordered_event_block_scan_over_many_blocks(semu):
    """The single-CTA scan of the block sums loops over chunks of 256 blocks (slabs above 512 Ki rows)."""
    rng = np.random.default_rng(5)
    n = 700
    sums = rng.integers(0, 5000, n).astype(u32)
    work = sums.copy()
    md = (O.EffectMetadata * 1)()
    md[0].base_child_index = 0
    child_infos = np.zeros((1, 2), dtype=np.int32)
    a = S.EventAppendArgs()
    a.block_sums, a.child_infos, a.metadata, a.binding = work.ctypes.data, child_infos.ctypes.data, C.addressof(md), 0
    semu.semu_events_scan_blocks(C.byref(a), n)
    np.testing.assert_array_equal(work, np.concatenate([[0], np.cumsum(sums)[:-1]]).astype(u32))
    assert child_infos[0, 1] == int(sums.sum())
