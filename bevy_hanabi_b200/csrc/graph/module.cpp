// module.cpp — expression module and its lowering to CUDA C. Restates reference src/graph/expr.rs:
// Module (:337-720), Expr::{is_const,has_side_effect,value_type,eval} (:996-1258), the literal /
// attribute / property / built-in expressions (:1270-1830) and the operator tables (:1830-2360), plus the
// ShaderWriter evaluation context of src/modifier/mod.rs:198-367.
//
// Lowering differences with the WGSL the reference emits (the expression STRUCTURE, parenthesisation
// and side-effect hoisting are identical):
//   f32 literal `1.`          -> `1.f`                      (C needs the suffix to stay in fp32)
//   `let varN = e;`           -> `const auto varN = e;`
//   `(a) % (b)`               -> `hnb_rem(a, b)`            (WGSL truncated remainder, also on floats)
//   `vec2(a, b)` etc.         -> `make_vec2(a, b)`          (C++ cannot overload a class template name)
//   `(a) * (b).x`             -> `((a) * (b)).x`            (swizzle of an infix operand: see Expr::Unary below)
#include "hanabi_graph.h"

namespace hnb_graph {

// ---- Module -------------------------------------------------------------------------------------
ExprHandle Module::add_expr(const Expr& e) {
    expressions_.push_back(e);
    return (ExprHandle)expressions_.size();
}
PropertyHandle Module::add_property(const std::string& name, const Value& default_value) {
    for (const auto& p : properties_)
        if (p.name == name) throw ExprError(ExprError::PropertyError, "duplicate property '" + name + "'");
    properties_.push_back({name, default_value});
    return (PropertyHandle)properties_.size();
}
const Property* Module::get_property(PropertyHandle h) const {
    if (h == 0 || h > properties_.size()) return nullptr;
    return &properties_[h - 1];
}
PropertyHandle Module::get_property_by_name(const std::string& name) const {
    for (size_t i = 0; i < properties_.size(); ++i)
        if (properties_[i].name == name) return (PropertyHandle)(i + 1);
    return 0;
}
void Module::gather_attributes(std::set<Attribute>& out) const {
    for (const auto& e : expressions_)
        if (e.kind == Expr::Attribute) out.insert(e.attribute);
}
const Expr& Module::try_get(ExprHandle h) const {
    if (h == 0 || h > expressions_.size())
        throw ExprError(ExprError::InvalidExprHandleError,
                        "Cannot find expression with handle " + std::to_string(h) +
                            " in the current module. Check that the Module used to build the expression was the same used in the EvalContext or the original EffectAsset.");
    return expressions_[h - 1];
}
ExprHandle Module::lit(const Value& v) {
    Expr e{};
    e.kind = Expr::Literal;
    e.literal = v;
    return add_expr(e);
}
ExprHandle Module::attr(Attribute a) {
    attribute_info(a);
    Expr e{};
    e.kind = Expr::Attribute;
    e.attribute = a;
    return add_expr(e);
}
ExprHandle Module::parent_attr(Attribute a) {
    attribute_info(a);
    Expr e{};
    e.kind = Expr::ParentAttribute;
    e.attribute = a;
    return add_expr(e);
}
ExprHandle Module::prop(PropertyHandle p) {
    Expr e{};
    e.kind = Expr::Property;
    e.property = p;
    return add_expr(e);
}
ExprHandle Module::builtin(BuiltInOperator op, ValueType rand_type) {
    if (op == BuiltInOperator::Rand && rand_type.is_matrix()) throw ExprError(ExprError::TypeError, "Invalid BuiltInOperator::Rand(ValueType::Matrix).");
    Expr e{};
    e.kind = Expr::BuiltIn;
    e.builtin = op;
    e.type = rand_type;
    return add_expr(e);
}
ExprHandle Module::unary(UnaryOperator op, ExprHandle x) {
    try_get(x);
    Expr e{};
    e.kind = Expr::Unary;
    e.op = (uint8_t)op;
    e.a = x;
    return add_expr(e);
}
ExprHandle Module::binary(BinaryOperator op, ExprHandle l, ExprHandle r) {
    try_get(l);
    try_get(r);
    Expr e{};
    e.kind = Expr::Binary;
    e.op = (uint8_t)op;
    e.a = l;
    e.b = r;
    return add_expr(e);
}
ExprHandle Module::ternary(TernaryOperator op, ExprHandle a, ExprHandle b, ExprHandle c) {
    try_get(a);
    try_get(b);
    try_get(c);
    Expr e{};
    e.kind = Expr::Ternary;
    e.op = (uint8_t)op;
    e.a = a;
    e.b = b;
    e.c = c;
    return add_expr(e);
}
ExprHandle Module::cast(ExprHandle x, ValueType target) {
    try_get(x);
    // CastExpr::is_valid (expr.rs:1468-1490): scalar->scalar, {scalar,vector}->vector, matrix->matrix
    if (auto t = value_type(x)) {
        bool ok = target.is_scalar() ? t->is_scalar() : target.is_vector() ? !t->is_matrix() : t->is_matrix();
        if (!ok) throw ExprError(ExprError::TypeError, "invalid cast from " + t->to_cuda_string() + " to " + target.to_cuda_string());
    }
    Expr e{};
    e.kind = Expr::Cast;
    e.a = x;
    e.type = target;
    return add_expr(e);
}

bool Module::is_const(ExprHandle h) const {
    const Expr& e = try_get(h);
    switch (e.kind) {
        case Expr::Literal: return true;
        case Expr::BuiltIn: case Expr::Property: case Expr::Attribute: case Expr::ParentAttribute: return false;
        case Expr::Unary: case Expr::Cast: return is_const(e.a);
        case Expr::Binary: return is_const(e.a) && is_const(e.b);
        case Expr::Ternary: return is_const(e.a) && is_const(e.b) && is_const(e.c);
    }
    return false;
}
bool Module::has_side_effect(ExprHandle h) const {
    const Expr& e = try_get(h);
    if (e.kind == Expr::BuiltIn) return e.builtin == BuiltInOperator::Rand;
    if (e.kind == Expr::Binary) return e.op == (uint8_t)BinaryOperator::UniformRand || e.op == (uint8_t)BinaryOperator::NormalRand;
    return false;
}
std::optional<ValueType> Module::value_type(ExprHandle h) const {
    const Expr& e = try_get(h);
    switch (e.kind) {
        case Expr::BuiltIn:
            if (e.builtin == BuiltInOperator::Rand) return e.type;
            if (e.builtin == BuiltInOperator::IsAlive) return BOOL;
            return FLOAT;
        case Expr::Literal: return e.literal.type;
        case Expr::Attribute: case Expr::ParentAttribute: return attribute_info(e.attribute).type;
        case Expr::Cast: return e.type;
        default: return std::nullopt;  // Property / Unary / Binary / Ternary: unknown, like the reference
    }
}

// ---- operator names -----------------------------------------------------------------------------
static const char* unary_name(UnaryOperator op) {
    static const char* n[] = {"abs", "acos", "asin", "atan", "all", "any", "ceil", "cos", "exp", "exp2", "floor", "fract",
                              "inverseSqrt", "length", "log", "log2", "normalize", "pack4x8snorm", "pack4x8unorm", "round",
                              "saturate", "sign", "sin", "sqrt", "tan", "unpack4x8snorm", "unpack4x8unorm", "w", "x", "y", "z"};
    return n[(int)op];
}
static bool unary_is_functional(UnaryOperator op) {
    return !(op == UnaryOperator::X || op == UnaryOperator::Y || op == UnaryOperator::Z || op == UnaryOperator::W);
}
static const char* binary_name(BinaryOperator op) {
    static const char* n[] = {"+", "atan2", "cross", "distance", "/", "dot", ">", ">=", "<", "<=", "max", "min", "*", "%", "step", "-",
                              "rand_uniform", "rand_normal", "make_vec2", "make_vec4"};
    return n[(int)op];
}
static bool binary_is_functional(BinaryOperator op) {
    switch (op) {
        case BinaryOperator::Add: case BinaryOperator::Div: case BinaryOperator::GreaterThan: case BinaryOperator::GreaterThanOrEqual:
        case BinaryOperator::LessThan: case BinaryOperator::LessThanOrEqual: case BinaryOperator::Mul: case BinaryOperator::Remainder:
        case BinaryOperator::Sub: return false;
        default: return true;
    }
}
static const char* ternary_name(TernaryOperator op) {
    static const char* n[] = {"mix", "clamp", "smoothstep", "make_vec3"};
    return n[(int)op];
}
static const char* builtin_name(BuiltInOperator op) {
    static const char* n[] = {"time", "delta_time", "virtual_time", "virtual_delta_time", "real_time", "real_delta_time", "rand", "alpha_cutoff", "is_alive"};
    return n[(int)op];
}

static std::string side_effect_local(ShaderWriter& ctx, const std::string& code, bool has_side_effect) {
    // check_side_effects_and_create_local_if_needed (expr.rs:1812-1824)
    if (!has_side_effect) return code;
    std::string var = ctx.make_local_var();
    ctx.push_stmt("const auto " + var + " = " + code + ";");
    return var;
}

std::string Module::eval_expr(ExprHandle h, ShaderWriter& ctx) const {
    const Expr& e = try_get(h);
    switch (e.kind) {
        case Expr::BuiltIn: {
            std::string code;
            if (e.builtin == BuiltInOperator::Rand) {
                // BuiltInOperator::name (expr.rs:1592-1628). Only the float generators exist in vfx_common.wgsl;
                // the reference would emit irand()/urand()/brand() which fail to compile — we reject them here.
                if (e.type.elem() != ScalarType::Float)
                    throw ExprError(ExprError::TypeError, "rand() is only available for f32 and vecN<f32> (vfx_common.wgsl defines no integer/bool generator)");
                int c = e.type.count();
                code = c == 1 ? "frand()" : "frand" + std::to_string(c) + "()";
            } else if (e.builtin == BuiltInOperator::IsAlive) {
                code = "is_alive";
            } else if (e.builtin == BuiltInOperator::AlphaCutoff) {
                throw ExprError(ExprError::GraphEvalError, "alpha_cutoff is a render-only built-in");
            } else {
                code = std::string("sim_params.") + builtin_name(e.builtin);
            }
            return side_effect_local(ctx, code, e.builtin == BuiltInOperator::Rand);
        }
        case Expr::Literal: return e.literal.to_cuda_string();
        case Expr::Property: {
            const hnb_graph::Property* p = get_property(e.property);
            if (!p) throw ExprError(ExprError::PropertyError, "Unknown property handle " + std::to_string(e.property) + " in evaluation module.");
            if (!ctx.property_layout.contains(p->name)) throw ExprError(ExprError::PropertyError, "Unknown property '" + p->name + "' in evaluation layout.");
            return "properties[properties_array_index]." + p->name;
        }
        case Expr::Attribute:
        case Expr::ParentAttribute: {
            const bool parent = e.kind == Expr::ParentAttribute;
            if (e.attribute == attr::ID) return parent ? "parent_particle_index" : "particle_index";
            if (e.attribute == attr::PARTICLE_COUNTER) return "particle_counter";
            const std::string owner = parent ? "parent_particle" : "particle";
            const std::string name = attribute_info(e.attribute).name;
            // the parent record is always a value; only the particle itself is a pointer inside modifier functions
            if (ctx.is_attribute_pointer() && !parent) return "(*" + owner + ")." + name;
            return owner + "." + name;
        }
        case Expr::Unary: {
            std::string inner = ctx.eval(*this, e.a);
            UnaryOperator op = (UnaryOperator)e.op;
            if (unary_is_functional(op)) return std::string(unary_name(op)) + "(" + inner + ")";
            // Deliberate deviation: the reference pastes ".x" behind an infix operand's text ("(a) * (b).x",
            // expr.rs:1146 with :1209), which binds the swizzle to the right operand only. The component of the
            // whole operand is what the node's value_type() declares, so parenthesise.
            const Expr& operand = expressions_[e.a - 1];
            if (operand.kind == Expr::Binary && !binary_is_functional((BinaryOperator)operand.op)) inner = "(" + inner + ")";
            return inner + "." + unary_name(op);
        }
        case Expr::Binary: {
            std::string l = ctx.eval(*this, e.a);
            std::string r = ctx.eval(*this, e.b);
            BinaryOperator op = (BinaryOperator)e.op;
            std::string body;
            if (binary_is_functional(op)) {
                if (op == BinaryOperator::UniformRand || op == BinaryOperator::NormalRand) {
                    auto lt = value_type(e.a), rt = value_type(e.b);
                    if (!lt || !rt) throw ExprError(ExprError::TypeError, "Can't determine the type of the operand");
                    if (*lt != *rt) throw ExprError(ExprError::TypeError, "Mismatched types");
                    std::string suffix;
                    if (*lt == FLOAT) suffix = "f";
                    else if (lt->is_vector() && lt->elem() == ScalarType::Float) suffix = "vec" + std::to_string(lt->count());
                    else throw ExprError(ExprError::TypeError, "Unsupported type");
                    body = std::string(binary_name(op)) + "_" + suffix + "(" + l + ", " + r + ")";
                } else {
                    body = std::string(binary_name(op)) + "(" + l + ", " + r + ")";
                }
            } else if (op == BinaryOperator::Remainder) {
                body = "hnb_rem(" + l + ", " + r + ")";
            } else {
                body = "(" + l + ") " + binary_name(op) + " (" + r + ")";
            }
            return side_effect_local(ctx, body, has_side_effect(h));
        }
        case Expr::Ternary: {
            std::string a = ctx.eval(*this, e.a);
            std::string b = ctx.eval(*this, e.b);
            std::string c = ctx.eval(*this, e.c);
            return std::string(ternary_name((TernaryOperator)e.op)) + "(" + a + ", " + b + ", " + c + ")";
        }
        case Expr::Cast: {
            std::string inner = ctx.eval(*this, e.a);
            return e.type.to_cuda_string() + "(" + inner + ")";
        }
    }
    throw ExprError(ExprError::GraphEvalError, "corrupt expression");
}

// ---- ShaderWriter -------------------------------------------------------------------------------
std::string ShaderWriter::eval(const Module& module, ExprHandle h) {
    auto it = expr_cache_.find(h);
    if (it != expr_cache_.end()) return it->second;
    std::string s = module.eval_expr(h, *this);
    expr_cache_[h] = s;
    return s;
}
std::string ShaderWriter::make_local_var() { return "var" + std::to_string(var_counter_++); }
void ShaderWriter::push_stmt(const std::string& stmt) {
    main_code += stmt;
    main_code += "\n";
}
void ShaderWriter::set_emits_gpu_spawn_events(bool use_events) {
    if (emits_.has_value() && *emits_ != use_events) throw ExprError(ExprError::GraphEvalError, "Conflicting use of GPU spawn events.");
    emits_ = use_events;
}

}  // namespace hnb_graph
