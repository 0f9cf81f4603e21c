"""Pins the CPU oracle (oracle/vfx_oracle.c) against every known-answer vector the reference's own tests
hold for the hot path:

  * src/render/headless_batching_tests.rs:74-107  / shader_contract_tests.rs:186-344   prefix-sum pass
  * src/render/headless_batching_tests.rs:110-150 / shader_contract_tests.rs:347-523   find_location_from_particle
  * src/render/shader_contract_tests.rs:636-884    indirect + routing (alive [7,5])
  * src/render/shader_contract_tests.rs:888-1230   the real generated update shader, 2 effects
  * src/render/shader_contract_tests.rs:1233-1489  the real indirect shader
  * src/render/mod.rs:7650-7724 (gpu_ops_ifda)     fill_dispatch_args
and the PRNG against hand-computed values of vfx_common.wgsl:266-335.

The same vectors are replayed against the CUDA kernels in tests/test_gpu_golden.py.
"""
import ctypes as C

import numpy as np

from oracle import c_oracle as O

u32p = C.POINTER(C.c_uint32)


def _p(a):
    return a.ctypes.data_as(u32p)


def test_prefix_sum_contract(orc):
    prefix = np.array([10, 5, 8, 6], dtype=np.uint32)
    batches = (O.BatchInfo * 2)()
    batches[0].spawner_base, batches[0].base_particle, batches[0].prefix_sum_offset, batches[0].prefix_sum_count = 0, 100, 0, 3
    batches[1].spawner_base, batches[1].base_particle, batches[1].prefix_sum_offset, batches[1].prefix_sum_count = 3, 500, 3, 1
    dispatch = np.zeros(6, dtype=np.uint32)
    orc.orc_prefix_sum(batches, 2, _p(prefix), _p(dispatch))
    assert prefix.tolist() == [0, 10, 15, 0]
    assert batches[0].total_update_count == 23
    assert batches[1].total_update_count == 6
    assert dispatch.tolist() == [1, 1, 1, 1, 1, 1]


def test_location_mapping(orc):
    prefix = np.array([0, 10, 15], dtype=np.uint32)
    bi = O.BatchInfo(0, 23, 7, 0, 0, 3)
    for idx, want in [(0, (0, 0, 0)), (10, (1, 10, 0)), (20, (2, 15, 5))]:
        loc = orc.orc_find_location_from_particle(C.byref(bi), _p(prefix), idx)
        assert (loc.effect_index, loc.base_particle, loc.update_index) == want
    # shader_contract_tests.rs:391-522: packed indices -> effect index, with the batch's particles at base 100
    got = []
    for packed in [0, 9, 10, 14, 15, 22]:
        got.append(orc.orc_find_location_from_particle(C.byref(bi), _p(prefix), packed).effect_index)
    assert got == [0, 0, 1, 1, 2, 2]


def test_prefix_sum_offset_is_honoured(orc):
    # a batch whose prefix entries do not start at 0 in the shared array
    prefix = np.array([99, 0, 4, 9], dtype=np.uint32)
    bi = O.BatchInfo(0, 0, 5, 0, 1, 3)
    loc = orc.orc_find_location_from_particle(C.byref(bi), _p(prefix), 8)
    assert (loc.effect_index, loc.base_particle, loc.update_index) == (1, 4, 4)


def test_real_indirect_contract(orc):
    sim = O.SimParams(1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 2)
    md = (O.EffectMetadata * 2)()
    md[0].capacity, md[0].alive_count, md[0].indirect_write_index, md[0].indirect_render_index = 200, 130, 0, 0
    md[1].capacity, md[1].alive_count, md[1].indirect_write_index, md[1].indirect_render_index = 5, 1, 1, 1
    draw = np.zeros(10, dtype=np.uint32)
    draw[1], draw[6] = 9, 4
    sp = (O.Spawner * 2)()
    sp[0].seed, sp[0].effect_metadata_index, sp[0].draw_indirect_index = 111, 0, 0
    sp[1].seed, sp[1].effect_metadata_index, sp[1].draw_indirect_index = 222, 1, 1
    prefix = np.zeros(2, dtype=np.uint32)
    orc.orc_indirect(C.byref(sim), md, _p(draw), sp, _p(prefix), None, 0)
    assert prefix.tolist() == [130, 1]
    assert (md[0].max_update, md[1].max_update) == (130, 1)
    assert (md[0].max_spawn, md[1].max_spawn) == (70, 4)
    assert (draw[1], draw[6]) == (0, 0)
    assert (md[0].indirect_write_index, md[1].indirect_write_index) == (1, 0)
    assert (sp[0].render_indirect_read_index, sp[1].render_indirect_read_index) == (1, 0)


def test_indirect_then_routing(orc):
    # shader_contract_tests.rs:636-884: alive [7,5] -> prefix [7,5] before the scan, max_update = alive; after the
    # scan 12 threads route to the two effects 7 / 5
    sim = O.SimParams(1.0, 0, 1.0, 0, 1.0, 0, 2)
    md = (O.EffectMetadata * 2)()
    md[0].capacity, md[0].alive_count, md[0].indirect_render_index = 16, 7, 0
    md[1].capacity, md[1].alive_count, md[1].indirect_render_index = 16, 5, 1
    draw = np.zeros(10, dtype=np.uint32)
    sp = (O.Spawner * 2)()
    sp[0].effect_metadata_index, sp[0].draw_indirect_index, sp[0].slab_offset = 0, 0, 0
    sp[1].effect_metadata_index, sp[1].draw_indirect_index, sp[1].slab_offset = 1, 1, 16
    prefix = np.zeros(2, dtype=np.uint32)
    orc.orc_indirect(C.byref(sim), md, _p(draw), sp, _p(prefix), None, 0)
    assert prefix.tolist() == [7, 5]
    assert (md[0].max_update, md[1].max_update) == (7, 5)
    batches = (O.BatchInfo * 1)(O.BatchInfo(0, 0, 0, 0, 0, 2))
    dispatch = np.zeros(3, dtype=np.uint32)
    orc.orc_prefix_sum(batches, 1, _p(prefix), _p(dispatch))
    assert prefix.tolist() == [0, 7] and batches[0].total_update_count == 12
    counts = [0, 0]
    for t in range(64):
        loc = orc.orc_find_location_from_particle(C.byref(batches[0]), _p(prefix), t)
        if loc.update_index < md[loc.effect_index].max_update:
            counts[loc.effect_index] += 1
    assert counts == [7, 5]


def test_real_update_contract(orc):
    # 2 effects in one slab of 8 rows (offsets 0 and 4), position-only layout (stride 4 u32), delta_time 1
    sim = O.SimParams(1.0, 0, 1.0, 0, 1.0, 0, 2)
    draw = np.zeros(10, dtype=np.uint32)
    particles = np.zeros((8, 4), dtype=np.uint32)
    indirect = np.zeros((8, 3), dtype=np.uint32)
    indirect[0, 0:2] = [0, 0]
    indirect[1, 0:2] = [0, 1]
    indirect[4, 0:2] = [0, 0]
    sp = (O.Spawner * 2)()
    sp[0].seed, sp[0].effect_metadata_index, sp[0].draw_indirect_index, sp[0].slab_offset, sp[0].parent_slab_offset = 1, 0, 0, 0, 0xFFFFFFFF
    sp[1].seed, sp[1].effect_metadata_index, sp[1].draw_indirect_index, sp[1].slab_offset, sp[1].parent_slab_offset = 2, 1, 1, 4, 0xFFFFFFFF
    md = (O.EffectMetadata * 2)()
    md[0].capacity, md[0].alive_count, md[0].max_update, md[0].indirect_render_index, md[0].particle_stride = 8, 2, 2, 0, 4
    md[1].capacity, md[1].alive_count, md[1].max_update, md[1].indirect_render_index, md[1].particle_stride = 8, 1, 1, 1, 4
    prefix = np.array([0, 2], dtype=np.uint32)
    bi = O.BatchInfo(0, 3, 0, 0, 0, 2)
    orc.orc_update(C.byref(sim), _p(draw), O.ptr(particles), 4, O.ptr(indirect), sp, _p(prefix), C.byref(bi), md, 64,
                   orc.orc_body_update_noop(), None)
    assert draw[1] == 2 and draw[6] == 1
    flat = indirect.reshape(-1)
    assert flat[0] == 0 and flat[3] == 1 and flat[12] == 0


def test_fill_dispatch_args(orc):
    # gpu_ops_ifda (mod.rs:7650-7724): thread counts -> ceil(n/64) workgroups, y = z = 1
    src = np.array([0, 1, 64, 65, 1000], dtype=np.uint32)
    dst = np.full(15, 7, dtype=np.uint32)
    orc.orc_fill_dispatch_args(_p(src), _p(dst), 0, 1, 0, 3, 5)
    assert dst.tolist() == [0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 16, 1, 1]
    # strided source (ChildInfo rows are 2 u32 wide; event_count is the second word)
    src2 = np.array([9, 130, 9, 5], dtype=np.uint32)
    dst2 = np.zeros(6, dtype=np.uint32)
    orc.orc_fill_dispatch_args(_p(src2), _p(dst2), 1, 2, 0, 3, 2)
    assert dst2.tolist() == [3, 1, 1, 1, 1, 1]


def test_pcg_hash_known_values(orc):
    # vfx_common.wgsl:266-270 evaluated by hand (python ints, mod 2^32)
    def ref(x):
        s = (x * 747796405 + 2891336453) & 0xFFFFFFFF
        w = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & 0xFFFFFFFF
        return (w >> 22) ^ w
    for x in [0, 1, 2, 42, 0xDEADBEEF, 0xFFFFFFFF, 123456789]:
        assert orc.orc_pcg_hash(x) == ref(x)
    assert orc.orc_to_float01(0) == 0.0
    assert 0.0 <= orc.orc_to_float01(0xFFFFFFFF) < 1.0
    # frand(): state advances once, value from a second hash (vfx_common.wgsl:278-281)
    seed = C.c_uint32(1234)
    v = orc.orc_frand(C.byref(seed))
    assert seed.value == ref(1234)
    import struct
    assert v == struct.unpack("<f", struct.pack("<I", (ref(ref(1234)) & 0x7FFFFF) | 0x3F800000))[0] - 1.0
    # frand4 packs 3 hashes into 4 lanes (:306-319)
    seed = C.c_uint32(77)
    out = (C.c_float * 4)()
    orc.orc_frand4(C.byref(seed), out)
    r0 = ref(77); r1 = ref(r0); r2 = ref(r1)
    assert seed.value == r2
    f01 = lambda u: struct.unpack("<f", struct.pack("<I", (u & 0x7FFFFF) | 0x3F800000))[0] - 1.0
    assert list(out) == [f01(r0), f01(((r0 & 0xFF000000) >> 8) | (r1 & 0xFFFF)), f01(((r1 & 0xFFFF0000) >> 8) | (r2 & 0xFF)), f01(r2 >> 8)]


def test_numpy_oracle_prng_matches_c(orc):
    from oracle import hanabi_oracle as H
    x = np.array([0, 1, 2, 42, 0xDEADBEEF, 0xFFFFFFFF], dtype=np.uint32)
    np.testing.assert_array_equal(H.pcg_hash(x), np.array([orc.orc_pcg_hash(int(v)) for v in x], dtype=np.uint32))
    rng = H.Rng(x.copy())
    for cnt in (1, 2, 3, 4):
        got = rng.frand_n(cnt)
    # replay in C
    seeds = [C.c_uint32(int(v)) for v in x]
    for cnt, fn in ((1, None), (2, orc.orc_frand2), (3, orc.orc_frand3), (4, orc.orc_frand4)):
        for s in seeds:
            if cnt == 1:
                last = [orc.orc_frand(C.byref(s))]
            else:
                buf = (C.c_float * 4)()
                fn(C.byref(s), buf)
                last = list(buf[:cnt])
    np.testing.assert_array_equal(rng.seed, np.array([s.value for s in seeds], dtype=np.uint32))
    np.testing.assert_array_equal(got[-1], np.array(last, dtype=np.float32))


def test_parallel_c5_update_equals_serial(orc):
    """The OpenMP baseline (orc_update_c5_parallel) must reproduce the serial thread-order oracle exactly."""
    from tests.helpers import Instance, RefWorld
    rng = np.random.default_rng(1)
    n = 5000
    def world():
        w = RefWorld(6000, 8, [Instance(0, 6000, alive=n, seed=9)])
        p = np.zeros((n, 8), dtype=np.float32)
        p[:, 0:3] = rng0.uniform(-1, 1, (n, 3)); p[:, 4:7] = rng0.uniform(-1, 1, (n, 3)); p[:, 7] = rng0.uniform(0.02, 0.3, n)
        w.particles[:n] = p.view(np.uint32)
        return w
    rng0 = np.random.default_rng(1); a = world()
    rng0 = np.random.default_rng(1); b = world()
    k = (C.c_float * 4)(0.0, -9.8, 0.0, 0.5)
    flags = np.zeros(6000, dtype=np.uint8)
    for step in range(12):
        a.oracle_frame(orc, orc.orc_body_update_c5(), k)
        b.oracle_indirect(orc); b.oracle_prefix_sum(orc)
        orc.orc_update_c5_parallel(C.byref(b.sim), _p(b.draw), O.ptr(b.particles), O.ptr(b.indirect), b.spawners, b.metadata, k, O.ptr(flags), 4)
        np.testing.assert_array_equal(a.particles, b.particles)
        np.testing.assert_array_equal(a.indirect, b.indirect)
        np.testing.assert_array_equal(a.metadata_rows(), b.metadata_rows())
        np.testing.assert_array_equal(a.draw, b.draw)
    assert a.metadata[0].alive_count < n
