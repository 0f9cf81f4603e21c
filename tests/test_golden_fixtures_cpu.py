"""Replays the committed fixtures of tests/golden/ (see its README): the reference's known-answer vectors against the
oracle and the host library, and the oracle's own recorded states against a fresh oracle run."""
import ctypes as C
import importlib.util
import json
from pathlib import Path

import numpy as np

from bevy_hanabi_b200 import graph as G
from oracle import c_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
KA = json.loads((GOLDEN / "reference_known_answers.json").read_text())
u32p = C.POINTER(C.c_uint32)


def _p(a):
    return a.ctypes.data_as(u32p)


def test_prefix_sum_vector(orc):
    v = KA["prefix_sum"]
    prefix = np.array(v["prefix_in"], dtype=np.uint32)
    batches = (O.BatchInfo * len(v["batches"]))()
    for b, row in zip(batches, v["batches"]):
        b.spawner_base, b.base_particle, b.prefix_sum_offset, b.prefix_sum_count = (row[k] for k in ("spawner_base", "base_particle", "prefix_sum_offset", "prefix_sum_count"))
    dispatch = np.zeros(3 * len(batches), dtype=np.uint32)
    orc.orc_prefix_sum(batches, len(batches), _p(prefix), _p(dispatch))
    assert prefix.tolist() == v["prefix_out"]
    assert [b.total_update_count for b in batches] == v["total_update_count"]
    assert dispatch.tolist() == v["dispatch_args"]


def test_location_mapping_vector(orc):
    v = KA["location_mapping"]
    prefix = np.array(v["prefix"], dtype=np.uint32)
    b = v["batch"]
    bi = O.BatchInfo(0, b["total_update_count"], b["spawner_base"], b["base_particle"], b["prefix_sum_offset"], b["prefix_sum_count"])
    for packed, want in v["packed_index_to_location"]:
        loc = orc.orc_find_location_from_particle(C.byref(bi), _p(prefix), packed)
        assert [loc.effect_index, loc.base_particle, loc.update_index] == want
    got = [orc.orc_find_location_from_particle(C.byref(bi), _p(prefix), x).effect_index for x in v["packed_indices"]]
    assert got == v["effect_indices"]


def test_real_indirect_vector(orc):
    v = KA["real_indirect"]
    n = len(v["metadata_in"])
    sim = O.SimParams(1.0, 0.0, 1.0, 0.0, 1.0, 0.0, n)
    md = (O.EffectMetadata * n)()
    sp = (O.Spawner * n)()
    draw = np.zeros(5 * n, dtype=np.uint32)
    for i, row in enumerate(v["metadata_in"]):
        for k, val in row.items():
            setattr(md[i], k, val)
        sp[i].effect_metadata_index = sp[i].draw_indirect_index = i
        draw[5 * i + 1] = v["instance_counts_in"][i]
    prefix = np.zeros(n, dtype=np.uint32)
    orc.orc_indirect(C.byref(sim), md, _p(draw), sp, _p(prefix), None, 0)
    assert prefix.tolist() == v["prefix_out"]
    assert [m.max_update for m in md] == v["max_update"]
    assert [m.max_spawn for m in md] == v["max_spawn"]
    assert [int(draw[5 * i + 1]) for i in range(n)] == v["instance_counts_out"]
    assert [m.indirect_write_index for m in md] == v["indirect_write_index_out"]
    assert [s.render_indirect_read_index for s in sp] == v["render_pong"]


def test_indirect_routing_vector(orc):
    v = KA["indirect_routing"]
    n = len(v["alive"])
    sim = O.SimParams(1.0, 0, 1.0, 0, 1.0, 0, n)
    md = (O.EffectMetadata * n)()
    sp = (O.Spawner * n)()
    for i, a in enumerate(v["alive"]):
        md[i].capacity, md[i].alive_count, md[i].indirect_render_index = v["capacity"], a, i
        sp[i].effect_metadata_index, sp[i].draw_indirect_index, sp[i].slab_offset = i, i, i * v["capacity"]
    draw = np.zeros(5 * n, dtype=np.uint32)
    prefix = np.zeros(n, dtype=np.uint32)
    orc.orc_indirect(C.byref(sim), md, _p(draw), sp, _p(prefix), None, 0)
    assert prefix.tolist() == v["prefix_after_indirect"]
    batches = (O.BatchInfo * 1)(O.BatchInfo(0, 0, 0, 0, 0, n))
    dispatch = np.zeros(3, dtype=np.uint32)
    orc.orc_prefix_sum(batches, 1, _p(prefix), _p(dispatch))
    assert prefix.tolist() == v["prefix_after_scan"] and batches[0].total_update_count == v["total_update_count"]
    counts = [0] * n
    for t in range(v["threads"]):
        loc = orc.orc_find_location_from_particle(C.byref(batches[0]), _p(prefix), t)
        if loc.update_index < md[loc.effect_index].max_update:
            counts[loc.effect_index] += 1
    assert counts == v["routed_counts"]


def test_real_update_vector(orc):
    v = KA["real_update"]
    n = len(v["instances"])
    sim = O.SimParams(v["delta_time"], 0, v["delta_time"], 0, v["delta_time"], 0, n)
    draw = np.zeros(5 * n, dtype=np.uint32)
    particles = np.zeros((v["slab_rows"], v["stride_words"]), dtype=np.uint32)
    indirect = np.zeros((v["slab_rows"], 3), dtype=np.uint32)
    for row, pp in v["indirect_rows"].items():
        indirect[int(row), 0:2] = pp
    sp = (O.Spawner * n)()
    md = (O.EffectMetadata * n)()
    for i, inst in enumerate(v["instances"]):
        sp[i].seed, sp[i].effect_metadata_index, sp[i].draw_indirect_index, sp[i].slab_offset, sp[i].parent_slab_offset = inst["seed"], i, i, inst["slab_offset"], 0xFFFFFFFF
        md[i].capacity, md[i].alive_count, md[i].max_update, md[i].indirect_render_index, md[i].particle_stride = v["slab_rows"], inst["alive_count"], inst["max_update"], i, v["stride_words"]
    prefix = np.array(v["prefix"], dtype=np.uint32)
    bi = O.BatchInfo(0, v["total_update_count"], 0, 0, 0, n)
    orc.orc_update(C.byref(sim), _p(draw), O.ptr(particles), v["stride_words"], O.ptr(indirect), sp, _p(prefix), C.byref(bi), md, v["threads"],
                   orc.orc_body_update_noop(), None)
    assert [int(draw[5 * i + 1]) for i in range(n)] == v["instance_counts"]
    flat = indirect.reshape(-1)
    for word, want in v["indirect_out_words"].items():
        assert flat[int(word)] == want


def test_fill_dispatch_vector(orc):
    v = KA["fill_dispatch_args"]
    src = np.array(v["thread_counts"], dtype=np.uint32)
    dst = np.zeros(3 * len(src), dtype=np.uint32)
    orc.orc_fill_dispatch_args(_p(src), _p(dst), 0, 1, 0, 3, len(src))
    assert dst.reshape(-1, 3)[:, 0].tolist() == v["workgroups"]
    assert dst.reshape(-1, 3)[:, 1:].tolist() == [[1, 1]] * len(src)


def test_host_side_vectors():
    for x, want in KA["f32_literals"]["cases"]:
        assert G.format_f32(x) == want + "f"
    for case in KA["particle_layouts"]["cases"]:
        fields, size, _ = G.particle_layout_of(case["attributes"])
        assert size == case["size"]
        assert [[f.offset, f.name] for f in fields] == case["fields"]
    w = G.ExprWriter()
    my_prop = w.add_property("my_prop", 3.0)
    x = w.lit(3.).abs().max(w.attr(G.Attribute.POSITION) * w.lit(2.)) + w.lit(-4.).min(w.prop(my_prop))
    text = w.finish().eval(x.expr())[0]
    assert text == KA["expression_text"]["cuda"]
    # the CUDA text is the reference's WGSL text with the f32 suffix on literals, nothing else
    assert text.replace(".f", ".") == KA["expression_text"]["wgsl"]


def test_slab_allocator_and_sorter_vectors():
    from bevy_hanabi_b200.cache import SliceAllocator
    from bevy_hanabi_b200.spawn import EffectSorter
    v = KA["slab_allocator"]
    a = SliceAllocator(v["capacity_request"])
    assert a.capacity == v["capacity"]
    slices = [a.allocate(n) for n in v["allocations"]]
    for k, used, nfree in zip(v["free_order"], v["used_size_after_each_free"], v["free_slice_count_after_each_free"]):
        a.free_slice(slices[k])
        assert (a.used_size, len(a.free_slices)) == (used, nfree)
    v = KA["effect_sorter"]
    s = EffectSorter()
    for entity, slab, base, parent in v["inserts"]:
        s.insert(entity, slab, base, parent)
    s.sort()
    assert s.entities() == v["sorted"]


def test_oracle_reproduces_its_recorded_states(orc):
    """Guards the oracle itself: the scenarios of make_oracle_fixtures.py must give the committed numbers."""
    spec = importlib.util.spec_from_file_location("make_oracle_fixtures", GOLDEN / "make_oracle_fixtures.py")
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    want = json.loads((GOLDEN / "oracle_states.json").read_text())
    assert mk.c5_with_deaths(orc) == want["c5_with_deaths"]
    assert mk.trails_bursts(orc) == want["trails_bursts"]
    assert 0 < want["c5_with_deaths"]["alive_per_step"][-1] < 4096
    assert want["trails_bursts"]["particle_counter"] == 3000
