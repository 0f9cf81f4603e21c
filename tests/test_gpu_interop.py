"""Device-resident interop: the render pass binds the particle buffer as AoS records and the indirect
buffer as interleaved rows on the device (vfx_render.wgsl:228-231, mod.rs:139-146). The exports must hand out exactly the
bytes the host download path returns — and the host path itself is pinned on the oracle's fill."""
import ctypes as C

import numpy as np
import pytest

from oracle import c_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stride", [4, 8, 12, 16, 20, 32, 36, 44, 48, 64, 100, 256])
@pytest.mark.parametrize("sector", [False, True])
def test_device_export_import_equals_host_path(ctx, stride, sector):
    rng = np.random.default_rng(stride * 2 + sector)
    rows = 5000                                  # not a multiple of the 256-row tile: the last CTA is partial
    slab = ctx.slab_create(rows, stride, sector_planes=sector)
    data = rng.integers(0, 2**32, (rows, stride // 4), dtype=np.uint32)
    ctx.slab_upload_aos(slab, 0, data)
    buf = ctx.device_alloc(rows * stride + 64)
    try:
        for first, count, skew in ((0, rows, 0), (37, 1234, 0), (4999, 1, 0), (100, 777, 4), (0, rows, 12)):
            ctx.slab_export_aos_device(slab, first, count, buf + skew)     # skew: destination not 16-byte aligned
            got = ctx.device_download(buf + skew, count * stride).reshape(count, stride // 4)
            np.testing.assert_array_equal(got, data[first:first + count])
        # import: overwrite a range from a device buffer, read everything back through the host path
        part = rng.integers(0, 2**32, (600, stride // 4), dtype=np.uint32)
        ctx.device_upload(buf, part)
        ctx.slab_import_aos_device(slab, 2000, 600, buf)
        data[2000:2600] = part
        np.testing.assert_array_equal(ctx.slab_download_aos(slab, 0, rows, stride), data)
    finally:
        ctx.device_free(buf)


def test_device_export_of_indirect_rows(ctx):
    rng = np.random.default_rng(5)
    rows = 3001
    slab = ctx.slab_create(rows, 32)
    ind = rng.integers(0, 2**32, (rows, 3), dtype=np.uint32)
    ctx.slab_upload_indirect(slab, 0, ind)
    buf = ctx.device_alloc(rows * 12 + 16)
    try:
        for first, count, skew in ((0, rows, 0), (5, 2000, 0), (1, 1023, 4), (3000, 1, 0)):
            ctx.slab_export_indirect_device(slab, first, count, buf + skew)
            got = ctx.device_download(buf + skew, count * 12).reshape(count, 3)
            np.testing.assert_array_equal(got, ind[first:first + count])
        new = rng.integers(0, 2**32, (500, 3), dtype=np.uint32)
        ctx.device_upload(buf, new)
        ctx.slab_import_indirect_device(slab, 1000, 500, buf)
        ind[1000:1500] = new
        np.testing.assert_array_equal(ctx.slab_download_indirect(slab, 0, rows), ind)
    finally:
        ctx.device_free(buf)


def test_device_view_describes_the_columns(ctx, orc):
    """The SoA columns in place: a renderer reading them directly sees the oracle's C5 state."""
    rows = 4096
    slab = ctx.slab_create(rows, 32)
    ctx.slab_fill_c5(slab, 0, rows, 99, 0.5, 2.0)
    v = ctx.slab_device_view(slab)
    assert (v.capacity_rows, v.particle_stride, v.num_planes) == (rows, 32, 2)
    assert list(v.plane_offset[:2]) == [0, 16] and list(v.plane_width[:2]) == [16, 16]
    assert v.planes[0] and v.planes[1] and v.ping and v.pong and v.dead
    ref = np.zeros((rows, 8), dtype=np.float32)
    ind = np.zeros((rows, 3), dtype=np.uint32)
    orc.orc_fill_c5(O.ptr(ref), O.ptr(ind), 0, rows, 99, 0.5, 2.0)
    for p in range(2):
        col = ctx.device_download(v.planes[p], rows * 16).reshape(rows, 4)
        np.testing.assert_array_equal(col, ref.view(np.uint32)[:, 4 * p:4 * p + 4])
    np.testing.assert_array_equal(ctx.device_download(v.ping, rows * 4), ind[:, 0])
    np.testing.assert_array_equal(ctx.device_download(v.dead, rows * 4), np.arange(rows, dtype=np.uint32))


def test_shard_fill_and_checksum_compose(ctx, orc):
    """hnb_slab_fill_c5_ex / hnb_slab_checksum_ex: two shards of a logical instance hold the unsharded values under local
    indices, and their checksums add up to the unsharded checksum (the oracle computes the same three numbers)."""
    P, cut, seed = 10000, 4321, 77
    whole = ctx.slab_create(P, 32)
    ctx.slab_fill_c5(whole, 0, P, seed, 0.1, 1.0)
    a, b = ctx.slab_create(cut, 32), ctx.slab_create(P - cut, 32)
    ctx.slab_fill_c5(a, 0, cut, seed, 0.1, 1.0, logical_first=0)
    ctx.slab_fill_c5(b, 0, P - cut, seed, 0.1, 1.0, logical_first=cut)
    full = ctx.slab_download_aos(whole, 0, P, 32)
    np.testing.assert_array_equal(ctx.slab_download_aos(a, 0, cut, 32), full[:cut])
    np.testing.assert_array_equal(ctx.slab_download_aos(b, 0, P - cut, 32), full[cut:])
    np.testing.assert_array_equal(ctx.slab_download_indirect(b, 0, P - cut)[:, 0], np.arange(P - cut, dtype=np.uint32))  # local indices
    cs = ctx.slab_checksum(whole, 0, P)
    assert (ctx.slab_checksum(a, 0, cut, index_base=0) + ctx.slab_checksum(b, 0, P - cut, index_base=cut)) % 2**64 == cs
    ref = np.zeros((P, 8), dtype=np.float32)
    orc.orc_fill_c5(O.ptr(ref), None, 0, P, seed, 0.1, 1.0)
    assert orc.orc_checksum(O.ptr(ref), 0, P, 8) == cs
    shard = np.zeros((P - cut, 8), dtype=np.float32)
    orc.orc_fill_c5_ex(O.ptr(shard), None, 0, P - cut, seed, 0.1, 1.0, cut)
    np.testing.assert_array_equal(shard, ref[cut:])
    assert orc.orc_checksum_ex(O.ptr(shard), 0, P - cut, 8, cut) == ctx.slab_checksum(b, 0, P - cut, index_base=cut)
