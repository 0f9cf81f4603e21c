// node_graph.cpp — the node-graph front end of the expression module (reference src/graph/node.rs):
//   * Graph: nodes, their slots, links between an output slot and input slots      (node.rs:244-443)
//   * Slot link rules: an output fans out, an input holds at most one source        (node.rs:200-240)
//   * Node::eval of AddNode / SubNode / MulNode / DivNode / AttributeNode / TimeNode / NormalizeNode: lowers the
//     node to expressions of a Module                                               (node.rs:446-775)
// plus one step the reference leaves to its caller: hnb_node_graph_eval_slot walks the links behind an output slot
// and evaluates the nodes in dependency order (memoised, cycle-checked), so that a wired graph lowers to the
// ExprHandle a modifier takes. Pure host code, exposed through include/hanabi_b200_graph.h.
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "hanabi_b200_graph.h"
#include "hanabi_graph.h"

using namespace hnb_graph;

extern "C" void hnb_set_last_error_(const char* msg);

namespace {

struct SlotDef {  // node.rs:82-141
    std::string name;
    bool is_input;
    int32_t value_type;  // -1 = None
};
struct Slot {  // node.rs:144-240
    uint32_t node_id, id;
    SlotDef def;
    std::vector<uint32_t> linked;
};
struct NodeRec {
    uint32_t kind;
    Attribute attr = 0;
    std::vector<SlotDef> defs;
};

const char* builtin_name(BuiltInOperator op) { return op == BuiltInOperator::Time ? "time" : "delta_time"; }  // expr.rs:1668-1672

std::vector<SlotDef> slot_defs(uint32_t kind, Attribute attr) {
    switch (kind) {
        case HNB_NODE_ADD:
        case HNB_NODE_SUB:
        case HNB_NODE_MUL:
        case HNB_NODE_DIV:  // node.rs:471-481
            return {{"lhs", true, -1}, {"rhs", true, -1}, {"result", false, -1}};
        case HNB_NODE_ATTRIBUTE:  // node.rs:653-658
            return {{attribute_info(attr).name, false, (int32_t)attribute_info(attr).type.code}};
        case HNB_NODE_TIME:  // node.rs:702-709
            return {{builtin_name(BuiltInOperator::Time), false, (int32_t)FLOAT.code}, {builtin_name(BuiltInOperator::DeltaTime), false, (int32_t)FLOAT.code}};
        case HNB_NODE_NORMALIZE:  // node.rs:741-747: both slots are declared as outputs in the reference
            return {{"in", false, -1}, {"out", false, -1}};
        default: throw ExprError(ExprError::GraphEvalError, "unknown node kind");
    }
}

// Node::eval (node.rs:488-775)
std::vector<ExprHandle> eval_node(const NodeRec& n, Module& m, const std::vector<ExprHandle>& in) {
    auto want = [&](size_t k, const char* who) {
        if (in.size() != k)
            throw ExprError(ExprError::GraphEvalError, std::string("Unexpected input count to ") + who + "::eval(): expected " +
                                                           std::to_string(k) + ", got " + std::to_string(in.size()));
    };
    switch (n.kind) {
        case HNB_NODE_ADD: want(2, "AddNode"); return {m.binary(BinaryOperator::Add, in[0], in[1])};
        case HNB_NODE_SUB: want(2, "SubNode"); return {m.binary(BinaryOperator::Sub, in[0], in[1])};
        case HNB_NODE_MUL: want(2, "MulNode"); return {m.binary(BinaryOperator::Mul, in[0], in[1])};
        case HNB_NODE_DIV: want(2, "DivNode"); return {m.binary(BinaryOperator::Div, in[0], in[1])};
        case HNB_NODE_ATTRIBUTE:
            if (!in.empty()) throw ExprError(ExprError::GraphEvalError, "Unexpected non-empty input to AttributeNode::eval().");
            return {m.attr(n.attr)};
        case HNB_NODE_TIME:
            if (!in.empty()) throw ExprError(ExprError::GraphEvalError, "Unexpected non-empty input to TimeNode::eval().");
            return {m.builtin(BuiltInOperator::Time), m.builtin(BuiltInOperator::DeltaTime)};
        case HNB_NODE_NORMALIZE:
            if (in.size() != 1) throw ExprError(ExprError::GraphEvalError, "Unexpected input slot count to NormalizeNode::eval() not equal to one.");
            return {m.unary(UnaryOperator::Normalize, in[0])};
        default: throw ExprError(ExprError::GraphEvalError, "unknown node kind");
    }
}

template <typename G> G& checked(G* g) {
    if (!g) throw ExprError(ExprError::GraphEvalError, "graph is NULL");
    return *g;
}

template <typename F> int32_t guarded(F&& f) {
    try {
        f();
        return HNB_OK;
    } catch (const std::exception& e) {
        hnb_set_last_error_(e.what());
        return HNB_ERR_EXPR;
    }
}

}  // namespace

struct hnb_node_graph {
    std::vector<NodeRec> nodes;
    std::vector<Slot> slots;
    Slot& slot(uint32_t id) {  // get_slot_mut (node.rs:438-442)
        if (id == 0 || id > slots.size()) throw ExprError(ExprError::GraphEvalError, "invalid slot id");
        return slots[id - 1];
    }
    const Slot& slot(uint32_t id) const { return const_cast<hnb_node_graph*>(this)->slot(id); }
    const NodeRec& node(uint32_t id) const {
        if (id == 0 || id > nodes.size()) throw ExprError(ExprError::GraphEvalError, "invalid node id");
        return nodes[id - 1];
    }
    // Evaluate the node that owns `node_id` once; inputs come from the slots linked to its input slots.
    const std::vector<ExprHandle>& eval(uint32_t node_id, Module& m, std::map<uint32_t, std::vector<ExprHandle>>& done,
                                        std::vector<uint8_t>& open) {
        auto it = done.find(node_id);
        if (it != done.end()) return it->second;
        const NodeRec& n = node(node_id);
        if (open[node_id - 1]) throw ExprError(ExprError::GraphEvalError, "the node graph has a cycle");
        open[node_id - 1] = 1;
        std::vector<ExprHandle> inputs;
        for (const Slot& s : slots) {
            if (s.node_id != node_id || !s.def.is_input) continue;
            if (s.linked.empty()) throw ExprError(ExprError::GraphEvalError, "input slot '" + s.def.name + "' is not linked");
            inputs.push_back(output_of(s.linked[0], m, done, open));
        }
        open[node_id - 1] = 0;
        return done.emplace(node_id, eval_node(n, m, inputs)).first->second;
    }
    ExprHandle output_of(uint32_t slot_id, Module& m, std::map<uint32_t, std::vector<ExprHandle>>& done, std::vector<uint8_t>& open) {
        const Slot& s = slot(slot_id);
        if (s.def.is_input) throw ExprError(ExprError::GraphEvalError, "slot '" + s.def.name + "' is not an output");
        const auto& outs = eval(s.node_id, m, done, open);
        uint32_t k = 0;  // index of this slot among the node's outputs
        for (const Slot& o : slots) {
            if (o.node_id != s.node_id || o.def.is_input) continue;
            if (o.id == slot_id) break;
            ++k;
        }
        if (k >= outs.size()) throw ExprError(ExprError::GraphEvalError, "slot '" + s.def.name + "' is not produced by its node's eval()");
        return outs[k];
    }
};

extern "C" {

hnb_node_graph* hnb_node_graph_create(void) { return new hnb_node_graph(); }
void hnb_node_graph_destroy(hnb_node_graph* g) { delete g; }

// Graph::add_node (node.rs:284-310): ids are 1-based; the node's slots get the next slot ids in declaration order
uint32_t hnb_node_graph_add_node(hnb_node_graph* g, uint32_t kind, const char* attribute) {
    uint32_t id = 0;
    guarded([&] {
        if (!g) throw ExprError(ExprError::GraphEvalError, "graph is NULL");
        NodeRec n;
        n.kind = kind;
        if (kind == HNB_NODE_ATTRIBUTE) {
            const Attribute a = attribute_by_name(attribute ? attribute : "position");  // AttributeNode::default = POSITION
            if (a < 0) throw ExprError(ExprError::GraphEvalError, std::string("unknown attribute '") + attribute + "'");
            n.attr = a;
        }
        n.defs = slot_defs(kind, n.attr);
        const uint32_t node_id = (uint32_t)g->nodes.size() + 1;
        for (const SlotDef& d : n.defs) g->slots.push_back(Slot{node_id, (uint32_t)g->slots.size() + 1, d, {}});
        g->nodes.push_back(std::move(n));
        id = node_id;
    });
    return id;
}
uint32_t hnb_node_graph_node_count(const hnb_node_graph* g) { return g ? (uint32_t)g->nodes.size() : 0; }

// Graph::link (node.rs:313-321); the reference asserts on the slot directions, here HNB_ERR_EXPR
int32_t hnb_node_graph_link(hnb_node_graph* g, uint32_t output, uint32_t input) {
    return guarded([&] {
        Slot& o = checked(g).slot(output);
        Slot& i = g->slot(input);
        if (o.def.is_input) throw ExprError(ExprError::GraphEvalError, "link: the first slot must be an output");
        if (!i.def.is_input) throw ExprError(ExprError::GraphEvalError, "link: the second slot must be an input");
        if (std::find(o.linked.begin(), o.linked.end(), input) == o.linked.end()) o.linked.push_back(input);  // link_to
        if (i.linked.empty()) i.linked.push_back(output); else i.linked[0] = output;                          // link_input
    });
}
// Graph::unlink (node.rs:330-338)
int32_t hnb_node_graph_unlink(hnb_node_graph* g, uint32_t output, uint32_t input) {
    return guarded([&] {
        Slot& o = checked(g).slot(output);
        Slot& i = g->slot(input);
        if (o.def.is_input) throw ExprError(ExprError::GraphEvalError, "unlink: the first slot must be an output");
        auto it = std::find(o.linked.begin(), o.linked.end(), input);
        if (it == o.linked.end()) return;
        o.linked.erase(it);
        if (!i.def.is_input) throw ExprError(ExprError::GraphEvalError, "unlink: the second slot must be an input");
        i.linked.clear();
    });
}
// Graph::unlink_all (node.rs:341-352)
int32_t hnb_node_graph_unlink_all(hnb_node_graph* g, uint32_t slot_id) {
    return guarded([&] {
        std::vector<uint32_t> linked;
        linked.swap(checked(g).slot(slot_id).linked);
        for (uint32_t r : linked) {
            Slot& remote = g->slot(r);
            if (remote.def.is_input) remote.linked.clear();
            else {
                auto it = std::find(remote.linked.begin(), remote.linked.end(), slot_id);
                if (it != remote.linked.end()) remote.linked.erase(it);
            }
        }
    });
}
// Graph::slots / input_slots / output_slots (node.rs:355-420): dir 0 = all, 1 = inputs, 2 = outputs
int32_t hnb_node_graph_slots(const hnb_node_graph* g, uint32_t node, uint32_t dir, uint32_t* out, uint32_t cap, uint32_t* n) {
    return guarded([&] {
        checked(g).node(node);
        uint32_t k = 0;
        for (const Slot& s : g->slots) {
            if (s.node_id != node) continue;
            if ((dir == 1 && !s.def.is_input) || (dir == 2 && s.def.is_input)) continue;
            if (out && k < cap) out[k] = s.id;
            ++k;
        }
        if (n) *n = k;
    });
}
// Graph::input_slot / output_slot (node.rs:369-407) and, with node == 0 and dir == 0, Graph::get_slot_id (:423-429):
// the first slot with that name, 0 if none
uint32_t hnb_node_graph_find_slot(const hnb_node_graph* g, uint32_t node, uint32_t dir, const char* name) {
    if (!g || !name) return 0;
    for (const Slot& s : g->slots) {
        if (node && s.node_id != node) continue;
        if ((dir == 1 && !s.def.is_input) || (dir == 2 && s.def.is_input)) continue;
        if (s.def.name == name) return s.id;
    }
    return 0;
}
int32_t hnb_node_graph_slot_info(const hnb_node_graph* g, uint32_t slot_id, const char** name, uint32_t* node, uint32_t* is_input,
                                 int32_t* value_type, uint32_t* linked, uint32_t cap, uint32_t* n_linked) {
    return guarded([&] {
        const Slot& s = checked(g).slot(slot_id);
        if (name) *name = s.def.name.c_str();
        if (node) *node = s.node_id;
        if (is_input) *is_input = s.def.is_input ? 1u : 0u;
        if (value_type) *value_type = s.def.value_type;
        for (uint32_t i = 0; linked && i < s.linked.size() && i < cap; ++i) linked[i] = s.linked[i];
        if (n_linked) *n_linked = (uint32_t)s.linked.size();
    });
}
// Node::eval (node.rs:458-463) of one node with explicit inputs
int32_t hnb_node_graph_eval_node(const hnb_node_graph* g, uint32_t node, hnb_module* m, const hnb_expr* inputs, uint32_t n_inputs,
                                 hnb_expr* outputs, uint32_t cap, uint32_t* n_outputs) {
    return guarded([&] {
        if (!m) throw ExprError(ExprError::GraphEvalError, "module is NULL");
        std::vector<ExprHandle> in(inputs, inputs + (inputs ? n_inputs : 0));
        const std::vector<ExprHandle> out = eval_node(checked(g).node(node), m->m, in);
        for (uint32_t i = 0; outputs && i < out.size() && i < cap; ++i) outputs[i] = out[i];
        if (n_outputs) *n_outputs = (uint32_t)out.size();
    });
}
// Lower the sub-graph behind an output slot: every node is evaluated once, inputs taken from the linked outputs.
int32_t hnb_node_graph_eval_slot(hnb_node_graph* g, hnb_module* m, uint32_t output_slot, hnb_expr* out) {
    return guarded([&] {
        if (!g || !m || !out) throw ExprError(ExprError::GraphEvalError, "NULL argument");
        std::map<uint32_t, std::vector<ExprHandle>> done;
        std::vector<uint8_t> open(g->nodes.size(), 0);
        *out = g->output_of(output_slot, m->m, done, open);
    });
}

}  // extern "C"
