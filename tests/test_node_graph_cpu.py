"""Node-graph front end (csrc/graph/node_graph.cpp) against the reference's own tests, src/graph/node.rs:781-968.
Expression text is compared in this backend's surface syntax (`3.f` for `3.`), like tests/test_authoring_cpu.py."""
import numpy as np
import pytest

from bevy_hanabi_b200 import graph as G
from bevy_hanabi_b200._native import HanabiError

A = G.Attribute


@pytest.mark.parametrize("make,sym", [(G.AddNode, "+"), (G.SubNode, "-"), (G.MulNode, "*"), (G.DivNode, "/")])
def test_binary_nodes(make, sym):
    """node.rs `add`, `sub`, `mul`, `div` (:781-876): wrong input counts are GraphEvalErrors, two inputs lower to
    `(lhs) op (rhs)`."""
    g, m = G.Graph(), G.Module()
    n = g.add_node(make())
    with pytest.raises(HanabiError, match="expected 2, got 0"):
        g.eval_node(n, m, [])
    three = m.lit(3.)
    with pytest.raises(HanabiError, match="expected 2, got 1"):
        g.eval_node(n, m, [three])
    two = m.lit(2.)
    outs = g.eval_node(n, m, [three, two])
    assert len(outs) == 1
    assert m.eval(outs[0])[0] == f"(3.f) {sym} (2.f)"
    s = [g.slot(i) for i in g.slots(n)]
    assert [(x.name, x.is_input, x.value_type) for x in s] == [("lhs", True, None), ("rhs", True, None), ("result", False, None)]


def test_attribute_node():
    """node.rs `attr` (:878-897)."""
    g, m = G.Graph(), G.Module()
    n = g.add_node(G.AttributeNode(A.POSITION))
    with pytest.raises(HanabiError, match="non-empty input"):
        g.eval_node(n, m, [m.lit(3.)])
    outs = g.eval_node(n, m, [])
    assert len(outs) == 1 and m.eval(outs[0])[0] == "particle.position"
    (sid,) = g.slots(n)
    assert g.slot(sid).name == "position" and g.slot(sid).value_type == G.VEC3 and not g.slot(sid).is_input
    assert g.slot(g.slots(g.add_node(G.AttributeNode()))[0]).name == "position"      # AttributeNode::default


def test_time_node():
    """node.rs `time` (:899-922)."""
    g, m = G.Graph(), G.Module()
    n = g.add_node(G.TimeNode())
    with pytest.raises(HanabiError, match="non-empty input"):
        g.eval_node(n, m, [m.lit(3.)])
    outs = g.eval_node(n, m, [])
    assert [m.eval(o)[0] for o in outs] == ["sim_params.time", "sim_params.delta_time"]
    assert [g.slot(s).name for s in g.output_slots(n)] == ["time", "delta_time"] and g.input_slots(n) == []


def test_normalize_node():
    """node.rs `normalize` (:924-942)."""
    g, m = G.Graph(), G.Module()
    n = g.add_node(G.NormalizeNode())
    with pytest.raises(HanabiError, match="not equal to one"):
        g.eval_node(n, m, [])
    outs = g.eval_node(n, m, [m.lit(G.Vec3(1, 1, 1))])
    assert len(outs) == 1 and m.eval(outs[0])[0] == "normalize(vec3<f32>(1.f,1.f,1.f))"


def _euler_graph():
    """node.rs `graph` (:944-968): position + velocity * delta_time."""
    g = G.Graph()
    nid_pos = g.add_node(G.AttributeNode(A.POSITION))
    nid_add = g.add_node(G.AddNode())
    sid_pos = g.output_slots(nid_pos)[0]
    sid_add_lhs, sid_add_rhs = g.input_slots(nid_add)
    g.link(sid_pos, sid_add_lhs)
    nid_vel = g.add_node(G.AttributeNode(A.VELOCITY))
    nid_mul = g.add_node(G.MulNode())
    nid_dt = g.add_node(G.TimeNode())
    sid_vel = g.output_slots(nid_vel)[0]
    sid_dt = g.output_slot(nid_dt, "delta_time")
    assert sid_dt is not None and g.output_slot(nid_dt, "nope") is None and g.input_slot(nid_dt, "delta_time") is None
    sid_mul_lhs, sid_mul_rhs = g.input_slots(nid_mul)
    g.link(sid_vel, sid_mul_lhs)
    g.link(sid_dt, sid_mul_rhs)
    sid_mul_out = g.output_slots(nid_mul)[0]
    g.link(sid_mul_out, sid_add_rhs)
    return g, dict(pos=sid_pos, add_lhs=sid_add_lhs, add_rhs=sid_add_rhs, vel=sid_vel, dt=sid_dt, mul_lhs=sid_mul_lhs,
                   mul_rhs=sid_mul_rhs, mul_out=sid_mul_out, add_out=g.output_slots(nid_add)[0], n_add=nid_add, n_mul=nid_mul)


def test_graph_topology():
    g, s = _euler_graph()
    # NodeId / SlotId numbering of Graph::add_node (node.rs:291-310): 1-based, slots in declaration order
    assert (s["pos"], s["add_lhs"], s["add_rhs"], s["add_out"]) == (1, 2, 3, 4)
    assert g.slot(s["pos"]).linked == [s["add_lhs"]] and g.slot(s["add_lhs"]).linked == [s["pos"]]
    assert g.slot(s["mul_out"]).linked == [s["add_rhs"]] and g.slot(s["add_rhs"]).linked == [s["mul_out"]]
    assert g.get_slot_id("result") == s["add_out"]                   # first slot of that name (node.rs:423-429)
    # direction checks (the reference asserts)
    with pytest.raises(HanabiError):
        g.link(s["add_lhs"], s["add_rhs"])
    with pytest.raises(HanabiError):
        g.link(s["pos"], s["vel"])
    # an input keeps one source: relinking replaces it (Slot::link_input, node.rs:222-229); the old output still lists it,
    # exactly as in the reference
    g.link(s["vel"], s["add_lhs"])
    assert g.slot(s["add_lhs"]).linked == [s["vel"]]
    assert set(g.slot(s["vel"]).linked) == {s["mul_lhs"], s["add_lhs"]}
    g.unlink(s["vel"], s["add_lhs"])                                 # Graph::unlink clears the input too
    assert g.slot(s["add_lhs"]).linked == [] and g.slot(s["vel"]).linked == [s["mul_lhs"]]
    g.unlink(s["vel"], s["add_lhs"])                                 # not linked: no-op
    g.unlink_all(s["mul_out"])                                       # Graph::unlink_all from the output side
    assert g.slot(s["mul_out"]).linked == [] and g.slot(s["add_rhs"]).linked == []
    g.link(s["dt"], s["mul_rhs"])
    g.unlink_all(s["mul_rhs"])                                       # ... and from the input side
    assert g.slot(s["mul_rhs"]).linked == [] and g.slot(s["dt"]).linked == []


def test_graph_lowers_to_one_expression():
    g, s = _euler_graph()
    m = G.Module()
    out = g.eval_slot(m, s["add_out"])
    assert m.eval(out)[0] == "(particle.position) + ((particle.velocity) * (sim_params.delta_time))"
    # the recorded mirror followed the native lowering: the expression is usable by a modifier and by the oracle
    assert [n.kind for n in m.nodes] == ["attr", "attr", "builtin", "builtin", "binary", "binary"]
    assert m.nodes[-1].op == "add" and m.nodes[-2].op == "mul"
    # unlinked input / cycle are reported
    g.unlink_all(s["mul_rhs"])
    with pytest.raises(HanabiError, match="'rhs' is not linked"):
        g.eval_slot(G.Module(), s["add_out"])
    g2 = G.Graph()
    a, b = g2.add_node(G.AddNode()), g2.add_node(G.AddNode())
    for n, other in ((a, b), (b, a)):
        for i in g2.input_slots(n):
            g2.link(g2.output_slots(other)[0], i)
    with pytest.raises(HanabiError, match="cycle"):
        g2.eval_slot(G.Module(), g2.output_slots(a)[0])


def test_graph_authored_effect_equals_the_hand_written_one():
    """A SetAttribute(POSITION, <graph>) update modifier generates the same code as the same expression written with
    the ExprWriter, and the numpy interpreter evaluates the adopted nodes."""
    g, s = _euler_graph()
    w = G.ExprWriter()
    zero = w.lit(G.Vec3(0, 0, 0))
    expr = g.eval_slot(w.module, s["add_out"])
    a1 = (G.EffectAsset(64, w.module, name="graph", motion_integration=G.MOTION_NONE)
          .init(G.SetAttributeModifier(A.POSITION, zero)).init(G.SetAttributeModifier(A.VELOCITY, w.lit(G.Vec3(1, 2, 3))))
          .update(G.SetAttributeModifier(A.POSITION, expr)))
    w2 = G.ExprWriter()
    zero2 = w2.lit(G.Vec3(0, 0, 0))
    expr2 = w2.attr(A.POSITION) + w2.attr(A.VELOCITY) * w2.delta_time()
    a2 = (G.EffectAsset(64, w2.module, name="graph", motion_integration=G.MOTION_NONE)
          .init(G.SetAttributeModifier(A.POSITION, zero2)).init(G.SetAttributeModifier(A.VELOCITY, w2.lit(G.Vec3(1, 2, 3))))
          .update(G.SetAttributeModifier(A.POSITION, expr2)))
    f1, f2 = a1.generate(), a2.generate()
    assert f1.update_code == f2.update_code and f1.init_code == f2.init_code
    assert "(particle.position) + ((particle.velocity) * (sim_params.delta_time))" in f1.update_code


def test_module_operator_methods_like_the_reference():
    """`impl_module_unary!` / `_binary!` / `_ternary!` (graph/expr.rs): one Module method per operator."""
    m = G.Module()
    a, b = m.lit(3.), m.lit(2.)
    assert m.eval(m.add(a, b))[0] == "(3.f) + (2.f)"
    assert m.eval(m.max(a, b))[0] == "max(3.f, 2.f)"
    assert m.eval(m.mix(a, b, m.lit(0.5)))[0] == "mix(3.f, 2.f, 0.5f)"
    v = m.lit(G.Vec3(1, 2, 3))
    assert m.eval(m.normalize(v))[0] == "normalize(vec3<f32>(1.f,2.f,3.f))"
    assert m.eval(m.y(v))[0] == "vec3<f32>(1.f,2.f,3.f).y"
    for op in G.UNARY + G.BINARY + G.TERNARY:
        assert callable(getattr(G.Module, op)), op
    p = m.add_property("speed", 4.0)
    assert m.get_property_by_name("speed") == p and m.get_property_by_name("nope") is None
    info = m.get(a)
    assert info.kind == 1 and info.value_type == G.FLOAT                 # Module::get: a literal
