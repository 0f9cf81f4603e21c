// hnb_wgsl.cuh — device prelude giving generated CUDA C the vocabulary the reference's generated
// WGSL relies on: vecN<T> types with component-wise operators, the WGSL built-in functions used by
// src/graph/expr.rs operator tables (:2035-2070, :2272-2296, :2349-2358), and the PRNG of
// src/render/vfx_common.wgsl:260-364 (bit-exact u32 arithmetic).
//
// Compiled by NVRTC (no system headers) and by nvcc. All float math is plain IEEE fp32: callers
// compile with -fmad=false and without fast-math so add/mul/div/sqrt sequences are bit-exact with the
// CPU oracle; only libm-style transcendentals may differ by a few ulp.
#pragma once

#ifndef HNB_SCALAR_TYPEDEFS
#define HNB_SCALAR_TYPEDEFS
typedef float f32;
typedef int i32;
typedef unsigned int u32;
typedef unsigned long long u64;
#endif

#define HNB_DI __device__ __forceinline__

// Everything lives in namespace hnb so that the WGSL built-in names (abs, min, max, sin, round, ...)
// hide CUDA's global overloads instead of colliding with them; generated code and the kernel
// templates are compiled inside the same namespace.
namespace hnb {

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+). A kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor in the stream still runs; hnb_pdl_wait() blocks until every predecessor grid
// has completed and its writes are visible (a no-op for a normally launched kernel), hnb_pdl_launch_dependents() lets
// the successor's CTAs take whatever SM slots free up from here on. Used on the frame chain
// update(N) -> bookkeeping(N+1) -> update(N+1): launch latency and the start skew of the persistent grid hide behind
// the predecessor's tail. (Compiled out on the host: the CPU kernel emulation of the test-suite.)
// ---------------------------------------------------------------------------------------------
HNB_DI void hnb_pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
HNB_DI void hnb_pdl_launch_dependents() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// Vector types
// ---------------------------------------------------------------------------------------------
template <typename T> struct vec2 {
    T x, y;
    HNB_DI vec2() : x(T(0)), y(T(0)) {}
    HNB_DI explicit vec2(T s) : x(s), y(s) {}
    HNB_DI vec2(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> HNB_DI explicit vec2(const vec2<U>& o) : x(T(o.x)), y(T(o.y)) {}
    HNB_DI T& operator[](int i) { return i == 0 ? x : y; }
    HNB_DI T operator[](int i) const { return i == 0 ? x : y; }
};
template <typename T> struct vec3 {
    T x, y, z;
    HNB_DI vec3() : x(T(0)), y(T(0)), z(T(0)) {}
    HNB_DI explicit vec3(T s) : x(s), y(s), z(s) {}
    HNB_DI vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    HNB_DI vec3(const vec2<T>& xy, T z_) : x(xy.x), y(xy.y), z(z_) {}
    HNB_DI vec3(T x_, const vec2<T>& yz) : x(x_), y(yz.x), z(yz.y) {}
    template <typename U> HNB_DI explicit vec3(const vec3<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    HNB_DI T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    HNB_DI T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T> struct vec4 {
    T x, y, z, w;
    HNB_DI vec4() : x(T(0)), y(T(0)), z(T(0)), w(T(0)) {}
    HNB_DI explicit vec4(T s) : x(s), y(s), z(s), w(s) {}
    HNB_DI vec4(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    HNB_DI vec4(const vec3<T>& xyz, T w_) : x(xyz.x), y(xyz.y), z(xyz.z), w(w_) {}
    HNB_DI vec4(const vec2<T>& xy, const vec2<T>& zw) : x(xy.x), y(xy.y), z(zw.x), w(zw.y) {}
    HNB_DI vec4(const vec2<T>& xy, T z_, T w_) : x(xy.x), y(xy.y), z(z_), w(w_) {}
    template <typename U> HNB_DI explicit vec4(const vec4<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
    HNB_DI T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    HNB_DI T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};
typedef vec2<f32> vec2f; typedef vec3<f32> vec3f; typedef vec4<f32> vec4f;
typedef vec2<i32> vec2i; typedef vec3<i32> vec3i; typedef vec4<i32> vec4i;
typedef vec2<u32> vec2u; typedef vec3<u32> vec3u; typedef vec4<u32> vec4u;

// WGSL swizzles used by generated code (emitted in functional form).
template <typename T> HNB_DI vec3<T> xyz(const vec4<T>& v) { return vec3<T>(v.x, v.y, v.z); }
template <typename T> HNB_DI vec2<T> xy(const vec4<T>& v) { return vec2<T>(v.x, v.y); }
template <typename T> HNB_DI vec2<T> xy(const vec3<T>& v) { return vec2<T>(v.x, v.y); }
template <typename T> HNB_DI vec3<T> xyz(const vec3<T>& v) { return v; }

// WGSL's type-inferring constructors `vec2(a, b)`, `vec3(a, b, c)`, `vec4(xyz, w)` (BinaryOperator::Vec2 /
// Vec4XyzW, TernaryOperator::Vec3 in src/graph/expr.rs): C++ cannot overload a class template's name with
// a function, so generated code calls these instead.
template <typename T> HNB_DI vec2<T> make_vec2(T x, T y) { return vec2<T>(x, y); }
template <typename T> HNB_DI vec3<T> make_vec3(T x, T y, T z) { return vec3<T>(x, y, z); }
template <typename T> HNB_DI vec4<T> make_vec4(const vec3<T>& xyz_, T w) { return vec4<T>(xyz_, w); }

// Apply a scalar function / operator component-wise.
#define HNB_VEC_UNARY(NAME, EXPR)                                                                     \
    template <typename T> HNB_DI vec2<T> NAME(const vec2<T>& a) { return vec2<T>(NAME(a.x), NAME(a.y)); } \
    template <typename T> HNB_DI vec3<T> NAME(const vec3<T>& a) { return vec3<T>(NAME(a.x), NAME(a.y), NAME(a.z)); } \
    template <typename T> HNB_DI vec4<T> NAME(const vec4<T>& a) { return vec4<T>(NAME(a.x), NAME(a.y), NAME(a.z), NAME(a.w)); }

#define HNB_VEC_BINARY_FN(NAME)                                                                       \
    template <typename T> HNB_DI vec2<T> NAME(const vec2<T>& a, const vec2<T>& b) { return vec2<T>(NAME(a.x, b.x), NAME(a.y, b.y)); } \
    template <typename T> HNB_DI vec3<T> NAME(const vec3<T>& a, const vec3<T>& b) { return vec3<T>(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z)); } \
    template <typename T> HNB_DI vec4<T> NAME(const vec4<T>& a, const vec4<T>& b) { return vec4<T>(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z), NAME(a.w, b.w)); }

#define HNB_VEC_ARITH(OP)                                                                             \
    template <typename T> HNB_DI vec2<T> operator OP(const vec2<T>& a, const vec2<T>& b) { return vec2<T>(a.x OP b.x, a.y OP b.y); } \
    template <typename T> HNB_DI vec3<T> operator OP(const vec3<T>& a, const vec3<T>& b) { return vec3<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
    template <typename T> HNB_DI vec4<T> operator OP(const vec4<T>& a, const vec4<T>& b) { return vec4<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
    template <typename T> HNB_DI vec2<T> operator OP(const vec2<T>& a, T b) { return vec2<T>(a.x OP b, a.y OP b); } \
    template <typename T> HNB_DI vec3<T> operator OP(const vec3<T>& a, T b) { return vec3<T>(a.x OP b, a.y OP b, a.z OP b); } \
    template <typename T> HNB_DI vec4<T> operator OP(const vec4<T>& a, T b) { return vec4<T>(a.x OP b, a.y OP b, a.z OP b, a.w OP b); } \
    template <typename T> HNB_DI vec2<T> operator OP(T a, const vec2<T>& b) { return vec2<T>(a OP b.x, a OP b.y); } \
    template <typename T> HNB_DI vec3<T> operator OP(T a, const vec3<T>& b) { return vec3<T>(a OP b.x, a OP b.y, a OP b.z); } \
    template <typename T> HNB_DI vec4<T> operator OP(T a, const vec4<T>& b) { return vec4<T>(a OP b.x, a OP b.y, a OP b.z, a OP b.w); } \
    template <typename T, typename U> HNB_DI vec2<T>& operator OP##=(vec2<T>& a, const U& b) { a = a OP b; return a; } \
    template <typename T, typename U> HNB_DI vec3<T>& operator OP##=(vec3<T>& a, const U& b) { a = a OP b; return a; } \
    template <typename T, typename U> HNB_DI vec4<T>& operator OP##=(vec4<T>& a, const U& b) { a = a OP b; return a; }

HNB_VEC_ARITH(+)
HNB_VEC_ARITH(-)
HNB_VEC_ARITH(*)
HNB_VEC_ARITH(/)

template <typename T> HNB_DI vec2<T> operator-(const vec2<T>& a) { return vec2<T>(-a.x, -a.y); }
template <typename T> HNB_DI vec3<T> operator-(const vec3<T>& a) { return vec3<T>(-a.x, -a.y, -a.z); }
template <typename T> HNB_DI vec4<T> operator-(const vec4<T>& a) { return vec4<T>(-a.x, -a.y, -a.z, -a.w); }

#define HNB_VEC_CMP(OP)                                                                               \
    template <typename T> HNB_DI vec2<bool> operator OP(const vec2<T>& a, const vec2<T>& b) { return vec2<bool>(a.x OP b.x, a.y OP b.y); } \
    template <typename T> HNB_DI vec3<bool> operator OP(const vec3<T>& a, const vec3<T>& b) { return vec3<bool>(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
    template <typename T> HNB_DI vec4<bool> operator OP(const vec4<T>& a, const vec4<T>& b) { return vec4<bool>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }
HNB_VEC_CMP(<)
HNB_VEC_CMP(<=)
HNB_VEC_CMP(>)
HNB_VEC_CMP(>=)

// WGSL `%`: truncated remainder. Integers use the C operator; floats e1 - e2 * trunc(e1 / e2).
HNB_DI f32 hnb_rem(f32 a, f32 b) { return a - b * truncf(a / b); }
HNB_DI i32 hnb_rem(i32 a, i32 b) { return a % b; }
HNB_DI u32 hnb_rem(u32 a, u32 b) { return a % b; }
HNB_VEC_BINARY_FN(hnb_rem)
template <typename T> HNB_DI vec2<T> hnb_rem(const vec2<T>& a, T b) { return vec2<T>(hnb_rem(a.x, b), hnb_rem(a.y, b)); }
template <typename T> HNB_DI vec3<T> hnb_rem(const vec3<T>& a, T b) { return vec3<T>(hnb_rem(a.x, b), hnb_rem(a.y, b), hnb_rem(a.z, b)); }
template <typename T> HNB_DI vec4<T> hnb_rem(const vec4<T>& a, T b) { return vec4<T>(hnb_rem(a.x, b), hnb_rem(a.y, b), hnb_rem(a.z, b), hnb_rem(a.w, b)); }

// ---------------------------------------------------------------------------------------------
// Scalar built-ins (WGSL names). `abs/min/max/...` on f32/i32/u32.
// ---------------------------------------------------------------------------------------------
HNB_DI f32 abs(f32 a) { return fabsf(a); }
HNB_DI f32 acos(f32 a) { return acosf(a); }
HNB_DI f32 asin(f32 a) { return asinf(a); }
HNB_DI f32 atan(f32 a) { return atanf(a); }
HNB_DI f32 ceil(f32 a) { return ceilf(a); }
HNB_DI f32 cos(f32 a) { return cosf(a); }
HNB_DI f32 exp(f32 a) { return expf(a); }
HNB_DI f32 exp2(f32 a) { return exp2f(a); }
HNB_DI f32 floor(f32 a) { return floorf(a); }
HNB_DI f32 fract(f32 a) { return a - floorf(a); }
HNB_DI f32 inverseSqrt(f32 a) { return 1.0f / sqrtf(a); }
HNB_DI f32 log(f32 a) { return logf(a); }
HNB_DI f32 log2(f32 a) { return log2f(a); }
HNB_DI f32 round(f32 a) { return rintf(a); }  // WGSL: ties to even
HNB_DI f32 sign(f32 a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
HNB_DI i32 sign(i32 a) { return a > 0 ? 1 : (a < 0 ? -1 : 0); }
HNB_DI f32 sin(f32 a) { return sinf(a); }
HNB_DI f32 sqrt(f32 a) { return sqrtf(a); }
HNB_DI f32 tan(f32 a) { return tanf(a); }
HNB_DI f32 min(f32 a, f32 b) { return fminf(a, b); }
HNB_DI f32 max(f32 a, f32 b) { return fmaxf(a, b); }
HNB_DI i32 min(i32 a, i32 b) { return a < b ? a : b; }
HNB_DI i32 max(i32 a, i32 b) { return a > b ? a : b; }
HNB_DI u32 min(u32 a, u32 b) { return a < b ? a : b; }
HNB_DI u32 max(u32 a, u32 b) { return a > b ? a : b; }
HNB_DI f32 atan2(f32 a, f32 b) { return atan2f(a, b); }
HNB_DI f32 pow(f32 a, f32 b) { return powf(a, b); }
HNB_DI f32 step(f32 edge, f32 x) { return edge <= x ? 1.0f : 0.0f; }
HNB_DI f32 clamp(f32 e, f32 lo, f32 hi) { return fminf(fmaxf(e, lo), hi); }
HNB_DI i32 clamp(i32 e, i32 lo, i32 hi) { return min(max(e, lo), hi); }
HNB_DI u32 clamp(u32 e, u32 lo, u32 hi) { return min(max(e, lo), hi); }
HNB_DI f32 saturate(f32 e) { return fminf(fmaxf(e, 0.0f), 1.0f); }
HNB_DI f32 mix(f32 a, f32 b, f32 t) { return a * (1.0f - t) + b * t; }
HNB_DI f32 smoothstep(f32 lo, f32 hi, f32 x) {
    f32 t = clamp((x - lo) / (hi - lo), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
HNB_DI f32 length(f32 a) { return fabsf(a); }
HNB_DI f32 distance(f32 a, f32 b) { return fabsf(a - b); }
HNB_DI i32 abs(i32 a) { return a < 0 ? -a : a; }
HNB_DI u32 abs(u32 a) { return a; }

HNB_VEC_UNARY(abs, ) HNB_VEC_UNARY(acos, ) HNB_VEC_UNARY(asin, ) HNB_VEC_UNARY(atan, ) HNB_VEC_UNARY(ceil, )
HNB_VEC_UNARY(cos, ) HNB_VEC_UNARY(exp, ) HNB_VEC_UNARY(exp2, ) HNB_VEC_UNARY(floor, ) HNB_VEC_UNARY(fract, )
HNB_VEC_UNARY(inverseSqrt, ) HNB_VEC_UNARY(log, ) HNB_VEC_UNARY(log2, ) HNB_VEC_UNARY(round, )
HNB_VEC_UNARY(sign, ) HNB_VEC_UNARY(sin, ) HNB_VEC_UNARY(sqrt, ) HNB_VEC_UNARY(tan, ) HNB_VEC_UNARY(saturate, )
HNB_VEC_BINARY_FN(min) HNB_VEC_BINARY_FN(max) HNB_VEC_BINARY_FN(atan2) HNB_VEC_BINARY_FN(pow) HNB_VEC_BINARY_FN(step)

#define HNB_VEC_TERNARY_FN(NAME)                                                                      \
    template <typename T> HNB_DI vec2<T> NAME(const vec2<T>& a, const vec2<T>& b, const vec2<T>& c) { return vec2<T>(NAME(a.x, b.x, c.x), NAME(a.y, b.y, c.y)); } \
    template <typename T> HNB_DI vec3<T> NAME(const vec3<T>& a, const vec3<T>& b, const vec3<T>& c) { return vec3<T>(NAME(a.x, b.x, c.x), NAME(a.y, b.y, c.y), NAME(a.z, b.z, c.z)); } \
    template <typename T> HNB_DI vec4<T> NAME(const vec4<T>& a, const vec4<T>& b, const vec4<T>& c) { return vec4<T>(NAME(a.x, b.x, c.x), NAME(a.y, b.y, c.y), NAME(a.z, b.z, c.z), NAME(a.w, b.w, c.w)); }
HNB_VEC_TERNARY_FN(mix) HNB_VEC_TERNARY_FN(clamp) HNB_VEC_TERNARY_FN(smoothstep)
// mix(vecN, vecN, f32) overload (WGSL allows a scalar blend factor)
template <typename T> HNB_DI vec2<T> mix(const vec2<T>& a, const vec2<T>& b, T t) { return vec2<T>(mix(a.x, b.x, t), mix(a.y, b.y, t)); }
template <typename T> HNB_DI vec3<T> mix(const vec3<T>& a, const vec3<T>& b, T t) { return vec3<T>(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
template <typename T> HNB_DI vec4<T> mix(const vec4<T>& a, const vec4<T>& b, T t) { return vec4<T>(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); }

// Geometric
template <typename T> HNB_DI T dot(const vec2<T>& a, const vec2<T>& b) { return a.x * b.x + a.y * b.y; }
template <typename T> HNB_DI T dot(const vec3<T>& a, const vec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> HNB_DI T dot(const vec4<T>& a, const vec4<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
HNB_DI vec3f cross(const vec3f& a, const vec3f& b) {
    return vec3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
HNB_DI f32 length(const vec2f& a) { return sqrtf(dot(a, a)); }
HNB_DI f32 length(const vec3f& a) { return sqrtf(dot(a, a)); }
HNB_DI f32 length(const vec4f& a) { return sqrtf(dot(a, a)); }
HNB_DI f32 normalize(f32 a) { return a / fabsf(a); }
HNB_DI vec2f normalize(const vec2f& a) { return a / length(a); }
HNB_DI vec3f normalize(const vec3f& a) { return a / length(a); }
HNB_DI vec4f normalize(const vec4f& a) { return a / length(a); }
HNB_DI f32 distance(const vec2f& a, const vec2f& b) { return length(a - b); }
HNB_DI f32 distance(const vec3f& a, const vec3f& b) { return length(a - b); }
HNB_DI f32 distance(const vec4f& a, const vec4f& b) { return length(a - b); }

// Logical reductions
HNB_DI bool all(bool a) { return a; }
HNB_DI bool any(bool a) { return a; }
HNB_DI bool all(const vec2<bool>& a) { return a.x && a.y; }
HNB_DI bool all(const vec3<bool>& a) { return a.x && a.y && a.z; }
HNB_DI bool all(const vec4<bool>& a) { return a.x && a.y && a.z && a.w; }
HNB_DI bool any(const vec2<bool>& a) { return a.x || a.y; }
HNB_DI bool any(const vec3<bool>& a) { return a.x || a.y || a.z; }
HNB_DI bool any(const vec4<bool>& a) { return a.x || a.y || a.z || a.w; }

// Packing (WGSL spec definitions)
HNB_DI u32 pack4x8unorm(const vec4f& v) {
    u32 r = 0;
    for (int i = 0; i < 4; ++i) r |= (u32(floorf(0.5f + 255.0f * fminf(fmaxf(v[i], 0.0f), 1.0f))) & 0xffu) << (8 * i);
    return r;
}
HNB_DI u32 pack4x8snorm(const vec4f& v) {
    u32 r = 0;
    for (int i = 0; i < 4; ++i) r |= (u32(i32(floorf(0.5f + 127.0f * fminf(fmaxf(v[i], -1.0f), 1.0f)))) & 0xffu) << (8 * i);
    return r;
}
HNB_DI vec4f unpack4x8unorm(u32 p) {
    return vec4f(f32(p & 0xffu) / 255.0f, f32((p >> 8) & 0xffu) / 255.0f, f32((p >> 16) & 0xffu) / 255.0f,
                 f32((p >> 24) & 0xffu) / 255.0f);
}
HNB_DI vec4f unpack4x8snorm(u32 p) {
    vec4f r;
    for (int i = 0; i < 4; ++i) {
        i32 b = i32(p << (24 - 8 * i)) >> 24;  // sign-extended byte
        r[i] = fmaxf(f32(b) / 127.0f, -1.0f);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// Matrices matCxR<f32>: C columns of R rows, column-major like WGSL, with WGSL's memory layout
// (array<vecR, C>: a vec3 column occupies 16 bytes) so that a matrix can sit in the Properties struct.
// Generated code builds them from C*R scalars (MatrixValue::to_wgsl_string, reference
// src/graph/mod.rs:1428-1441) and combines them with the WGSL arithmetic operators: m + m, m - m,
// m * s, s * m, m * v (vecC -> vecR), v * m (vecR -> vecC), m * m. Sums run over the columns from
// left to right, one rounding per operation.
// ---------------------------------------------------------------------------------------------
template <int R> struct hnb_col;
template <> struct hnb_col<2> { typedef vec2f vec; vec2f v; };
template <> struct hnb_col<3> { typedef vec3f vec; vec3f v; f32 _pad; };
template <> struct hnb_col<4> { typedef vec4f vec; vec4f v; };

template <int C, int R> struct hnb_mat {
    typedef typename hnb_col<R>::vec col_t;
    hnb_col<R> c[C];
    HNB_DI hnb_mat() {}
    template <typename... A> HNB_DI explicit hnb_mat(f32 e0, A... rest) {
        static_assert(sizeof...(A) + 1 == C * R, "a matCxR literal takes C*R scalars");
        const f32 e[C * R] = {e0, f32(rest)...};
#pragma unroll
        for (int j = 0; j < C; ++j)
#pragma unroll
            for (int i = 0; i < R; ++i) c[j].v[i] = e[j * R + i];
    }
    HNB_DI col_t& operator[](int j) { return c[j].v; }
    HNB_DI const col_t& operator[](int j) const { return c[j].v; }
};
// The 4x4 case keeps bare vec4 columns: the kernel templates build the spawner transforms through `c[]`.
template <> struct hnb_mat<4, 4> {
    typedef vec4f col_t;
    vec4f c[4];
    HNB_DI hnb_mat() {}
    HNB_DI hnb_mat(f32 e0, f32 e1, f32 e2, f32 e3, f32 e4, f32 e5, f32 e6, f32 e7, f32 e8, f32 e9, f32 e10, f32 e11, f32 e12, f32 e13,
                   f32 e14, f32 e15) {
        c[0] = vec4f(e0, e1, e2, e3);
        c[1] = vec4f(e4, e5, e6, e7);
        c[2] = vec4f(e8, e9, e10, e11);
        c[3] = vec4f(e12, e13, e14, e15);
    }
    HNB_DI vec4f& operator[](int i) { return c[i]; }
    HNB_DI const vec4f& operator[](int i) const { return c[i]; }
};
typedef hnb_mat<2, 2> mat2x2f; typedef hnb_mat<2, 3> mat2x3f; typedef hnb_mat<2, 4> mat2x4f;
typedef hnb_mat<3, 2> mat3x2f; typedef hnb_mat<3, 3> mat3x3f; typedef hnb_mat<3, 4> mat3x4f;
typedef hnb_mat<4, 2> mat4x2f; typedef hnb_mat<4, 3> mat4x3f; typedef hnb_mat<4, 4> mat4x4f;
static_assert(sizeof(mat2x2f) == 16 && sizeof(mat3x2f) == 24 && sizeof(mat4x2f) == 32, "matCx2: 8-byte columns");
static_assert(sizeof(mat2x3f) == 32 && sizeof(mat3x3f) == 48 && sizeof(mat4x3f) == 64, "matCx3: 16-byte columns");
static_assert(sizeof(mat2x4f) == 32 && sizeof(mat3x4f) == 48 && sizeof(mat4x4f) == 64, "matCx4: 16-byte columns");

HNB_DI vec4f operator*(const mat4x4f& m, const vec4f& v) {
    // WGSL: sum over columns, accumulated left to right
    return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w;
}
template <int C, int R> HNB_DI typename hnb_col<R>::vec operator*(const hnb_mat<C, R>& m, const typename hnb_col<C>::vec& v) {
    typename hnb_col<R>::vec acc = m[0] * v[0];
#pragma unroll
    for (int j = 1; j < C; ++j) acc = acc + m[j] * v[j];
    return acc;
}
template <int C, int R> HNB_DI typename hnb_col<C>::vec operator*(const typename hnb_col<R>::vec& v, const hnb_mat<C, R>& m) {
    typename hnb_col<C>::vec r;
#pragma unroll
    for (int j = 0; j < C; ++j) r[j] = dot(v, m[j]);
    return r;
}
template <int K, int R, int C> HNB_DI hnb_mat<C, R> operator*(const hnb_mat<K, R>& a, const hnb_mat<C, K>& b) {
    hnb_mat<C, R> r;
#pragma unroll
    for (int j = 0; j < C; ++j) r[j] = a * b[j];
    return r;
}
template <int C, int R> HNB_DI hnb_mat<C, R> operator*(const hnb_mat<C, R>& a, f32 s) {
    hnb_mat<C, R> r;
#pragma unroll
    for (int j = 0; j < C; ++j) r[j] = a[j] * s;
    return r;
}
template <int C, int R> HNB_DI hnb_mat<C, R> operator*(f32 s, const hnb_mat<C, R>& a) {
    hnb_mat<C, R> r;
#pragma unroll
    for (int j = 0; j < C; ++j) r[j] = s * a[j];
    return r;
}
template <int C, int R> HNB_DI hnb_mat<C, R> operator+(const hnb_mat<C, R>& a, const hnb_mat<C, R>& b) {
    hnb_mat<C, R> r;
#pragma unroll
    for (int j = 0; j < C; ++j) r[j] = a[j] + b[j];
    return r;
}
template <int C, int R> HNB_DI hnb_mat<C, R> operator-(const hnb_mat<C, R>& a, const hnb_mat<C, R>& b) {
    hnb_mat<C, R> r;
#pragma unroll
    for (int j = 0; j < C; ++j) r[j] = a[j] - b[j];
    return r;
}
// transpose(mat4x4(row0,row1,row2,(0,0,0,1))) as built in vfx_init.wgsl:157-164
HNB_DI mat4x4f hnb_transform_from_rows(const f32* r0, const f32* r1, const f32* r2) {
    mat4x4f m;
    m.c[0] = vec4f(r0[0], r1[0], r2[0], 0.0f);
    m.c[1] = vec4f(r0[1], r1[1], r2[1], 0.0f);
    m.c[2] = vec4f(r0[2], r1[2], r2[2], 0.0f);
    m.c[3] = vec4f(r0[3], r1[3], r2[3], 1.0f);
    return m;
}

// ---------------------------------------------------------------------------------------------
// PRNG (reference src/render/vfx_common.wgsl:260-364). `seed` is per-thread state passed by
// reference; generated code reaches it through the frand*() macros defined by the kernel template.
// ---------------------------------------------------------------------------------------------
#define HNB_TAU 6.283185307179586476925286766559f
constexpr f32 tau = HNB_TAU;  // vfx_common.wgsl:262

HNB_DI u32 pcg_hash(u32 input) {
    u32 state = input * 747796405u + 2891336453u;
    u32 word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
HNB_DI f32 to_float01(u32 u) { return __uint_as_float((u & 0x007fffffu) | 0x3f800000u) - 1.0f; }
HNB_DI f32 hnb_frand(u32& seed) {
    seed = pcg_hash(seed);
    return to_float01(pcg_hash(seed));
}
HNB_DI vec2f hnb_frand2(u32& seed) {
    seed = pcg_hash(seed); f32 x = to_float01(seed);
    seed = pcg_hash(seed); f32 y = to_float01(seed);
    return vec2f(x, y);
}
HNB_DI vec3f hnb_frand3(u32& seed) {
    seed = pcg_hash(seed); f32 x = to_float01(seed);
    seed = pcg_hash(seed); f32 y = to_float01(seed);
    seed = pcg_hash(seed); f32 z = to_float01(seed);
    return vec3f(x, y, z);
}
HNB_DI vec4f hnb_frand4(u32& seed) {
    u32 r0 = pcg_hash(seed);
    u32 r1 = pcg_hash(r0);
    u32 r2 = pcg_hash(r1);
    seed = r2;
    f32 x = to_float01(r0);
    u32 r01 = (r0 & 0xff000000u) >> 8u | (r1 & 0x0000ffffu);
    f32 y = to_float01(r01);
    u32 r12 = (r1 & 0xffff0000u) >> 8u | (r2 & 0x000000ffu);
    f32 z = to_float01(r12);
    u32 r22 = r2 >> 8u;
    f32 w = to_float01(r22);
    return vec4f(x, y, z, w);
}
HNB_DI f32 hnb_rand_uniform_f(u32& seed, f32 a, f32 b) { return a + hnb_frand(seed) * (b - a); }
HNB_DI vec2f hnb_rand_uniform_vec2(u32& seed, const vec2f& a, const vec2f& b) { return a + hnb_frand2(seed) * (b - a); }
HNB_DI vec3f hnb_rand_uniform_vec3(u32& seed, const vec3f& a, const vec3f& b) { return a + hnb_frand3(seed) * (b - a); }
HNB_DI vec4f hnb_rand_uniform_vec4(u32& seed, const vec4f& a, const vec4f& b) { return a + hnb_frand4(seed) * (b - a); }
HNB_DI f32 hnb_rand_normal_f(u32& seed, f32 mean, f32 std_dev) {
    f32 u = hnb_frand(seed);
    f32 v = hnb_frand(seed);
    f32 r = sqrtf(-2.0f * logf(u));
    return mean + std_dev * r * cosf(HNB_TAU * v);
}
HNB_DI vec2f hnb_rand_normal_vec2(u32& seed, const vec2f& mean, const vec2f& std_dev) {
    f32 u = hnb_frand(seed);
    vec2f v = hnb_frand2(seed);
    f32 r = sqrtf(-2.0f * logf(u));
    return mean + std_dev * r * cos(HNB_TAU * v);
}
HNB_DI vec3f hnb_rand_normal_vec3(u32& seed, const vec3f& mean, const vec3f& std_dev) {
    f32 u = hnb_frand(seed);
    vec3f v = hnb_frand3(seed);
    f32 r = sqrtf(-2.0f * logf(u));
    return mean + std_dev * r * cos(HNB_TAU * v);
}
HNB_DI vec4f hnb_rand_normal_vec4(u32& seed, const vec4f& mean, const vec4f& std_dev) {
    f32 u = hnb_frand(seed);
    vec4f v = hnb_frand4(seed);
    f32 r = sqrtf(-2.0f * logf(u));
    return mean + std_dev * r * cos(HNB_TAU * v);
}
HNB_DI vec3f proj(const vec3f& u, const vec3f& v) { return dot(v, u) / dot(u, u) * u; }

// Names generated code uses (the WGSL `var<private> seed` becomes a local reference named `seed`).
#define frand() hnb_frand(seed)
#define frand2() hnb_frand2(seed)
#define frand3() hnb_frand3(seed)
#define frand4() hnb_frand4(seed)
#define rand_uniform_f(a, b) hnb_rand_uniform_f(seed, a, b)
#define rand_uniform_vec2(a, b) hnb_rand_uniform_vec2(seed, a, b)
#define rand_uniform_vec3(a, b) hnb_rand_uniform_vec3(seed, a, b)
#define rand_uniform_vec4(a, b) hnb_rand_uniform_vec4(seed, a, b)
#define rand_normal_f(a, b) hnb_rand_normal_f(seed, a, b)
#define rand_normal_vec2(a, b) hnb_rand_normal_vec2(seed, a, b)
#define rand_normal_vec3(a, b) hnb_rand_normal_vec3(seed, a, b)
#define rand_normal_vec4(a, b) hnb_rand_normal_vec4(seed, a, b)

}  // namespace hnb
