// value_attr.cpp — value types, literal formatting, the attribute table, ParticleLayout and
// PropertyLayout. Restates reference src/lib.rs:264-430 (ToWgslString), src/attributes.rs:152-675
// (types, attributes, defaults), :1516-1670 (ParticleLayoutBuilder::build), src/properties.rs:437-453
// (serialize), :561-699 (PropertyLayout::new).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "hanabi_graph.h"

namespace hnb_graph {

// ---- ValueType ----------------------------------------------------------------------------------
ValueType ValueType::vector(ScalarType s, int count) {
    if (count < 2 || count > 4) throw ExprError(ExprError::TypeError, "invalid vector size");
    switch (s) {
        case ScalarType::Bool: return ValueType(4u + (uint32_t)(count - 2));
        case ScalarType::Float: return ValueType(7u + (uint32_t)(count - 2));
        case ScalarType::Int: return ValueType(10u + (uint32_t)(count - 2));
        default: return ValueType(13u + (uint32_t)(count - 2));
    }
}
ScalarType ValueType::elem() const {
    if (code < 4) return (ScalarType)code;
    if (code < 7) return ScalarType::Bool;
    if (code < 10) return ScalarType::Float;
    if (code < 13) return ScalarType::Int;
    if (code < 16) return ScalarType::Uint;
    return ScalarType::Float;
}
// Matrix codes: 16..18 = mat2x2, mat3x3, mat4x4; 19..24 = mat2x3, mat2x4, mat3x2, mat3x4, mat4x2, mat4x3 (CxR).
static const uint8_t kMatCols[9] = {2, 3, 4, 2, 2, 3, 3, 4, 4};
static const uint8_t kMatRows[9] = {2, 3, 4, 3, 4, 2, 4, 2, 3};
ValueType ValueType::matrix(int cols, int rows) {
    if (cols < 2 || cols > 4 || rows < 2 || rows > 4) throw ExprError(ExprError::TypeError, "invalid matrix size");
    for (uint32_t i = 0; i < 9; ++i)
        if (kMatCols[i] == cols && kMatRows[i] == rows) return ValueType(16u + i);
    return ValueType(16u);
}
int ValueType::cols() const { return is_matrix() && is_valid() ? kMatCols[code - 16] : 1; }
int ValueType::rows() const { return is_matrix() && is_valid() ? kMatRows[code - 16] : count(); }
int ValueType::count() const {
    if (code < 4) return 1;
    if (code < 16) return 2 + (int)((code - 4) % 3);
    return cols() * (is_valid() ? kMatRows[code - 16] : 1);
}
uint32_t ValueType::size() const {
    if (!is_matrix()) return 4u * (uint32_t)count();
    return (uint32_t)cols() * (rows() >= 3 ? 16u : 8u);
}
uint32_t ValueType::align() const {
    if (is_scalar()) return 4;
    if (is_vector()) return count() == 2 ? 8 : 16;
    return rows() == 2 ? 8 : 16;
}
static const char* scalar_name(ScalarType s) {
    switch (s) {
        case ScalarType::Bool: return "bool";
        case ScalarType::Float: return "f32";
        case ScalarType::Int: return "i32";
        default: return "u32";
    }
}
std::string ValueType::to_cuda_string() const {
    if (is_scalar()) return scalar_name(elem());
    if (is_vector()) return "vec" + std::to_string(count()) + "<" + scalar_name(elem()) + ">";
    return "mat" + std::to_string(cols()) + "x" + std::to_string(rows()) + "f";
}

// ---- Value --------------------------------------------------------------------------------------
Value Value::from_f32(float f) {
    Value v;
    v.type = FLOAT;
    memcpy(&v.bits[0], &f, 4);
    return v;
}
Value Value::from_words(ValueType t, const uint32_t* words) {
    Value v;
    v.type = t;
    for (int i = 0; i < t.count(); ++i) v.bits[i] = words[i];
    return v;
}
float Value::f(int i) const {
    float x;
    memcpy(&x, &bits[i], 4);
    return x;
}

std::string f32_to_cuda_string(float f) {
    if (std::isnan(f)) return "__int_as_float(0x7fc00000)";
    if (std::isinf(f)) return f > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
    char buf[64];
    snprintf(buf, sizeof(buf), "%.6f", (double)f);  // exact binary value rounded to 6 decimals, like Rust's {:.6}
    std::string s(buf);
    while (!s.empty() && s.back() == '0') s.pop_back();  // "1.000000" -> "1."
    return s + "f";
}

static std::string scalar_to_string(ScalarType t, uint32_t bits) {
    switch (t) {
        case ScalarType::Bool: return bits ? "true" : "false";
        case ScalarType::Float: {
            float f;
            memcpy(&f, &bits, 4);
            return f32_to_cuda_string(f);
        }
        case ScalarType::Int: return std::to_string((int32_t)bits);
        default: return std::to_string(bits) + "u";
    }
}

std::string Value::to_cuda_string() const {
    if (type.is_scalar()) return scalar_to_string(type.elem(), bits[0]);
    // matrices: every component in storage order = column by column (MatrixValue::to_wgsl_string, graph/mod.rs:1428-1441)
    std::string s = type.to_cuda_string() + "(";
    for (int i = 0; i < type.count(); ++i) {
        if (i) s += ",";
        s += scalar_to_string(type.elem(), bits[i]);
    }
    return s + ")";
}

// ---- Attributes ---------------------------------------------------------------------------------
namespace {
Value vf(std::initializer_list<float> xs) {
    Value v;
    v.type = xs.size() == 1 ? FLOAT : ValueType::vector(ScalarType::Float, (int)xs.size());
    int i = 0;
    for (float x : xs) memcpy(&v.bits[i++], &x, 4);
    return v;
}
Value vu(uint32_t x) {
    Value v;
    v.type = UINT;
    v.bits[0] = x;
    return v;
}
Value vi(int32_t x) {
    Value v;
    v.type = INT;
    v.bits[0] = (uint32_t)x;
    return v;
}
const std::vector<AttributeInfo>& table() {
    static const std::vector<AttributeInfo> t = [] {
        std::vector<AttributeInfo> a;
        auto add = [&](const char* n, const Value& d) { a.push_back({n, d.type, d}); };
        add("id", vu(0));
        add("particle_counter", vu(0));
        add("position", vf({0, 0, 0}));
        add("velocity", vf({0, 0, 0}));
        add("age", vf({0}));
        add("lifetime", vf({1}));
        add("color", vu(0xFFFFFFFFu));
        add("hdr_color", vf({1, 1, 1, 1}));
        add("alpha", vf({1}));
        add("size", vf({1}));
        add("size2", vf({1, 1}));
        add("size3", vf({1, 1, 1}));
        add("prev", vu(0xFFFFFFFFu));
        add("next", vu(0xFFFFFFFFu));
        add("axis_x", vf({1, 0, 0}));
        add("axis_y", vf({0, 1, 0}));
        add("axis_z", vf({0, 0, 1}));
        add("sprite_index", vi(0));
        add("f32_0", vf({0})); add("f32_1", vf({0})); add("f32_2", vf({0})); add("f32_3", vf({0}));
        add("f32x2_0", vf({0, 0})); add("f32x2_1", vf({0, 0})); add("f32x2_2", vf({0, 0})); add("f32x2_3", vf({0, 0}));
        add("f32x3_0", vf({0, 0, 0})); add("f32x3_1", vf({0, 0, 0})); add("f32x3_2", vf({0, 0, 0})); add("f32x3_3", vf({0, 0, 0}));
        add("f32x4_0", vf({0, 0, 0, 0})); add("f32x4_1", vf({0, 0, 0, 0})); add("f32x4_2", vf({0, 0, 0, 0})); add("f32x4_3", vf({0, 0, 0, 0}));
        add("u32_0", vu(0)); add("u32_1", vu(0)); add("u32_2", vu(0)); add("u32_3", vu(0));
        add("ribbon_id", vu(0));
        return a;
    }();
    return t;
}
}  // namespace

int attribute_count() { return (int)table().size(); }
const AttributeInfo& attribute_info(Attribute a) {
    if (a < 0 || a >= attribute_count()) throw ExprError(ExprError::GraphEvalError, "invalid attribute id");
    return table()[(size_t)a];
}
Attribute attribute_by_name(const std::string& name) {
    for (int i = 0; i < attribute_count(); ++i)
        if (name == table()[(size_t)i].name) return i;
    return -1;
}
namespace attr {
const Attribute ID = 0, PARTICLE_COUNTER = 1, POSITION = 2, VELOCITY = 3, AGE = 4, LIFETIME = 5, COLOR = 6, HDR_COLOR = 7, ALPHA = 8,
                SIZE = 9, SIZE2 = 10, SIZE3 = 11, PREV = 12, NEXT = 13, AXIS_X = 14, AXIS_Y = 15, AXIS_Z = 16, SPRITE_INDEX = 17,
                RIBBON_ID = 38;
}

// ---- ParticleLayout -----------------------------------------------------------------------------
ParticleLayout ParticleLayout::build(const std::set<Attribute>& attrs_in) {
    struct Item {
        Attribute a;
        std::string name;
        uint32_t size;
        ValueType type;
    };
    std::vector<Item> items;
    for (Attribute a : attrs_in) {
        const auto& info = attribute_info(a);
        items.push_back({a, info.name, info.type.size(), info.type});
    }
    // "Remove duplicates" sorts by name, then "Sort by size". The reference uses sort_unstable_by_key; for the
    // handful of attributes of a layout Rust's small-sort is insertion based, i.e. order inside a size class stays
    // alphabetical. We use a stable sort to pin exactly that order.
    std::sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.name < y.name; });
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.size < y.size; });

    ParticleLayout out;
    uint32_t offset = 0, align = 4;
    int next_pad = 0;
    auto push = [&](const Item& it, uint32_t off) { out.layout.push_back({it.name, it.type, off, false}); };
    auto push_pad = [&](uint32_t off) {
        if (next_pad >= 5) throw ExprError(ExprError::GraphEvalError, "particle layout needs too many padding fields");
        out.layout.push_back({"pad" + std::to_string(next_pad++), UINT, off, true});
    };
    auto part = [&](uint32_t sz) { return (size_t)(std::partition_point(items.begin(), items.end(), [&](const Item& i) { return i.size < sz; }) - items.begin()); };
    const size_t n = items.size();
    // all Float4
    size_t index4 = part(16);
    for (size_t i = index4; i < n; ++i) { push(items[i], offset); offset += 16; }
    if (n - index4 > 0) align = 16;
    size_t index2 = part(8);
    size_t num1 = index2;
    size_t index3 = part(12);
    size_t num2 = index3 - index2;
    size_t num3 = index4 - index3;
    if (num3 > 0) align = 16;
    else if (num2 > 0) align = std::max<uint32_t>(align, 8);
    // paired { Float3 + Float1 }
    size_t num_pairs = std::min(num1, num3);
    for (size_t i = 0; i < num_pairs; ++i) {
        push(items[index3 + i], offset); offset += 12;
        push(items[i], offset); offset += 4;
    }
    size_t index1 = num_pairs;
    index3 += num_pairs;
    num1 -= num_pairs;
    num3 -= num_pairs;
    // paired { Float2 + Float2 }
    for (size_t i = 0; i < num2 / 2; ++i)
        for (size_t j = 0; j < 2; ++j) { push(items[index2 + i * 2 + j], offset); offset += 8; }
    index2 += (num2 / 2) * 2;
    num2 %= 2;
    // remaining Float3, each padded to 16 bytes
    for (size_t i = 0; i < num3; ++i) {
        push(items[index3 + i], offset);
        push_pad(offset + 12);
        offset += 16;
    }
    // the single Float2 if any
    if (num2 > 0) { push(items[index2], offset); offset += 8; }
    // remaining Float1
    for (size_t i = 0; i < num1; ++i) { push(items[index1 + i], offset); offset += 4; }
    // pad the struct to its alignment (wgpu issue 5262 workaround in the reference)
    uint32_t padded = (offset + align - 1) / align * align;
    while (offset < padded) { push_pad(offset); offset += 4; }
    out.align = align;
    return out;
}

uint32_t ParticleLayout::size() const {
    if (layout.empty()) return 0;
    const auto& last = layout.back();
    return last.offset + last.type.size();
}
bool ParticleLayout::contains(Attribute a) const {
    const char* n = attribute_info(a).name;
    for (const auto& l : layout)
        if (!l.is_pad && l.name == n) return true;
    return false;
}
std::optional<uint32_t> ParticleLayout::byte_offset(Attribute a) const {
    const char* n = attribute_info(a).name;
    for (const auto& l : layout)
        if (!l.is_pad && l.name == n) return l.offset;
    return std::nullopt;
}

// ---- PropertyLayout -----------------------------------------------------------------------------
PropertyLayout PropertyLayout::make(const std::vector<Property>& props_in) {
    std::vector<const Property*> props;
    for (const auto& p : props_in) props.push_back(&p);
    // sort_unstable_by_key(size) in the reference; stable here (insertion order inside a size class), see above
    std::stable_sort(props.begin(), props.end(), [](const Property* x, const Property* y) { return x->default_value.type.size() < y->default_value.type.size(); });
    auto size_of = [&](size_t i) { return props[i]->default_value.type.size(); };
    auto part = [&](uint32_t sz) {
        size_t i = 0;
        while (i < props.size() && size_of(i) < sz) ++i;
        return i;
    };
    PropertyLayout out;
    uint32_t offset = 0;
    auto push = [&](size_t i, uint32_t off) { out.layout.push_back({*props[i], off}); };
    const size_t n = props.size();
    size_t index4 = part(16);
    for (size_t i = index4; i < n; ++i) { push(i, offset); offset += 16; }
    size_t index2 = part(8), num1 = index2, index3 = part(12);
    size_t num2 = index3 - index2, num3 = index4 - index3;
    size_t num_pairs = std::min(num1, num3);
    for (size_t i = 0; i < num_pairs; ++i) {
        push(index3 + i, offset); offset += 12;
        push(i, offset); offset += 4;
    }
    size_t index1 = num_pairs;
    index3 += num_pairs;
    num1 -= num_pairs;
    num3 -= num_pairs;
    for (size_t i = 0; i < num2 / 2; ++i)
        for (size_t j = 0; j < 2; ++j) { push(index2 + i * 2 + j, offset); offset += 8; }
    index2 += (num2 / 2) * 2;
    num2 %= 2;
    if (num3 > num1) {
        for (size_t i = 0; i < num3; ++i) { push(index3 + i, offset); offset += 16; }
        if (num2 > 0) push(index2, offset);
    } else {
        if (num2 > 0) { push(index2, offset); offset += 8; }
        for (size_t i = 0; i < num1; ++i) { push(index1 + i, offset); offset += 4; }
    }
    return out;
}
uint32_t PropertyLayout::cpu_size() const {
    if (layout.empty()) return 0;
    const auto& last = layout.back();
    return last.offset + last.property.default_value.type.size();
}
uint32_t PropertyLayout::align() const {
    uint32_t a = 0;
    for (const auto& e : layout) a = std::max(a, e.property.default_value.type.align());
    return a;
}
uint32_t PropertyLayout::min_binding_size() const {
    if (layout.empty()) return 0;
    uint32_t a = align();
    return (cpu_size() + a - 1) / a * a;
}
bool PropertyLayout::contains(const std::string& name) const { return offset(name).has_value(); }
std::optional<uint32_t> PropertyLayout::offset(const std::string& name) const {
    for (const auto& e : layout)
        if (e.property.name == name) return e.offset;
    return std::nullopt;
}
std::string PropertyLayout::generate_struct_body() const {
    std::string s;
    uint32_t cursor = 0;
    int pad = 0;
    for (const auto& e : layout) {
        // PropertyLayout::new advances by 16 bytes after every property of 16 bytes or more (properties.rs:572-580):
        // a matrix larger than that is only laid out consistently when nothing follows it. The reference would
        // silently overlap the next field; refuse instead.
        if (cursor > e.offset)
            throw ExprError(ExprError::Validate, "property '" + e.property.name + "' overlaps the matrix property before it (the reference's property layout only holds a matrix larger than 16 bytes as its last entry: the largest property, with no scalar, vec2 or vec3 property beside it)");
        while (cursor < e.offset) { s += "    u32 _hnb_pad" + std::to_string(pad++) + ";\n"; cursor += 4; }
        ValueType t = e.property.default_value.type;
        if (t.elem() == ScalarType::Bool) {
            // bools are stored as 32-bit words (0 / 0xFFFFFFFF), graph/mod.rs:103-109
            t = t.is_scalar() ? UINT : ValueType::vector(ScalarType::Uint, t.count());
        }
        s += "    " + t.to_cuda_string() + " " + e.property.name + ";\n";
        cursor = e.offset + t.size();
    }
    uint32_t total = min_binding_size();
    while (cursor < total) { s += "    u32 _hnb_pad" + std::to_string(pad++) + ";\n"; cursor += 4; }
    return s;
}
std::vector<uint8_t> PropertyLayout::serialize(const std::vector<std::pair<std::string, Value>>& values) const {
    std::vector<uint8_t> data(min_binding_size(), 0);
    for (const auto& e : layout) {
        const Value* v = &e.property.default_value;
        for (const auto& kv : values)
            if (kv.first == e.property.name) v = &kv.second;
        if (v->type != e.property.default_value.type) throw ExprError(ExprError::PropertyError, "property '" + e.property.name + "' set with a value of the wrong type");
        memcpy(&data[e.offset], v->bits, v->type.size());
    }
    return data;
}

}  // namespace hnb_graph
