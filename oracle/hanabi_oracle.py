"""hanabi_oracle.py — numpy restatement of the reference's simulation passes for ARBITRARY effects.
TEST INFRASTRUCTURE ONLY: only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

Where ``vfx_oracle.c`` restates the pass structure thread by thread for fixed effect bodies, this module
interprets an effect authored through ``bevy_hanabi_b200.graph`` — the recorded expression nodes and
modifier list, NOT the generated CUDA code — so that it checks the expression compiler as well as the
kernels. It follows, independently of the product's C++ lowering:

  expressions   src/graph/expr.rs: operator semantics = the WGSL built-ins the reference emits
                (:2035-2070, :2272-2296, :2349-2358); side-effect (rand) expressions are evaluated once
                per writer at first use (:1812-1824, modifier/mod.rs:309-319), pure ones re-read the
                particle at every use like the emitted text does
  PRNG          src/render/vfx_common.wgsl:266-364
  modifiers     src/modifier/{accel,force,kill,attr,position,velocity}.rs  (formulas in SURVEY.md App. B)
  aging/reaping src/lib.rs:1223-1264, Euler integration :1106-1121, global-space translation :525-528
  init pass     src/render/vfx_init.wgsl:101-196     update pass  src/render/vfx_update.wgsl:106-167
  indirect      src/render/vfx_indirect.wgsl:31-90   prefix sum   src/render/vfx_prefix_sum.wgsl:14-43

All instances' threads are executed "in ascending global_invocation_id" — vectorised per instance, which
is equivalent because no thread reads another particle's record — and the alive / dead lists are written
in that canonical order (stable compaction).

fp32 discipline: every arithmetic step is a separate numpy float32 operation (no FMA, no float64
intermediates), dot products are summed left to right. Transcendentals (sin, cos, acos, pow, log, sqrt is
exact) come from numpy's libm and may differ from CUDA's by a few ulp: parity on those paths is checked
with the 1e-5 relative tolerance of BASELINE.json, everything else bit-exactly.

Parity status: "parity unpinned" for floating-point modifier results (no reference test executes them,
SURVEY.md §8c); the integer bookkeeping agrees with vfx_oracle.c, which is pinned by the reference's
known-answer tests (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
U32 = np.uint32
TAU = F32(6.283185307179586476925286766559)


# ---------------------------------------------------------------------------------------------------
# PRNG (vfx_common.wgsl:266-364), vectorised over threads
# ---------------------------------------------------------------------------------------------------
def pcg_hash(x: np.ndarray) -> np.ndarray:
    x = x.astype(U32)
    with np.errstate(over="ignore"):
        state = x * U32(747796405) + U32(2891336453)
        word = ((state >> ((state >> U32(28)) + U32(4))) ^ state) * U32(277803737)
    return (word >> U32(22)) ^ word


def to_float01(u: np.ndarray) -> np.ndarray:
    bits = (u.astype(U32) & U32(0x007FFFFF)) | U32(0x3F800000)
    return bits.view(F32) - F32(1.0)


class Rng:
    """The `var<private> seed` of a dispatch: one u32 per thread."""

    def __init__(self, seed: np.ndarray):
        self.seed = seed.astype(U32)

    def frand(self):
        self.seed = pcg_hash(self.seed)
        return to_float01(pcg_hash(self.seed))

    def frand_n(self, n: int):
        if n == 1:
            return self.frand()
        if n in (2, 3):
            out = []
            for _ in range(n):
                self.seed = pcg_hash(self.seed)
                out.append(to_float01(self.seed))
            return np.stack(out, axis=1)
        r0 = pcg_hash(self.seed)
        r1 = pcg_hash(r0)
        r2 = pcg_hash(r1)
        self.seed = r2
        x = to_float01(r0)
        y = to_float01(((r0 & U32(0xFF000000)) >> U32(8)) | (r1 & U32(0x0000FFFF)))
        z = to_float01(((r1 & U32(0xFFFF0000)) >> U32(8)) | (r2 & U32(0x000000FF)))
        w = to_float01(r2 >> U32(8))
        return np.stack([x, y, z, w], axis=1)


# ---------------------------------------------------------------------------------------------------
# value helpers: scalars are (n,) arrays, vectors (n, c) arrays
# ---------------------------------------------------------------------------------------------------
def _bc(a, b):
    """Broadcast scalar (n,) against vector (n,c)."""
    if a.ndim == 1 and b.ndim == 2:
        return a[:, None], b
    if a.ndim == 2 and b.ndim == 1:
        return a, b[:, None]
    return a, b


def dot(a, b):
    p = a * b
    s = p[:, 0]
    for i in range(1, p.shape[1]):
        s = s + p[:, i]
    return s


def length(a):
    return np.abs(a) if a.ndim == 1 else np.sqrt(dot(a, a))


def normalize(a):
    l = length(a)
    return a / l if a.ndim == 1 else a / l[:, None]


def cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def mix(a, b, t):
    a, t1 = _bc(a, t)
    b, t2 = _bc(b, t)
    return a * (F32(1.0) - t1) + b * t2


def clamp(e, lo, hi):
    return np.minimum(np.maximum(e, lo), hi)


def smoothstep(lo, hi, x):
    t = clamp((x - lo) / (hi - lo), F32(0.0), F32(1.0))
    return t * t * (F32(3.0) - F32(2.0) * t)


def sign(a):
    return np.where(a > 0, F32(1.0), np.where(a < 0, F32(-1.0), F32(0.0))).astype(F32) if a.dtype == F32 else np.sign(a)


def step(edge, x):
    edge, x = _bc(edge, x)
    return np.where(edge <= x, F32(1.0), F32(0.0)).astype(F32)


def _trunc_rem(a, b):
    if a.dtype == F32:
        return a - b * np.trunc(a / b)
    return np.fmod(a, b)


_DT = {"f": F32, "u": U32, "i": np.int32, "b": np.bool_}


# matCxR<f32> value-type codes -> (columns, rows) (MatrixType, reference src/attributes.rs:322-397)
_MAT_DIMS = {16: (2, 2), 17: (3, 3), 18: (4, 4), 19: (2, 3), 20: (2, 4), 21: (3, 2), 22: (3, 4), 23: (4, 2), 24: (4, 3)}


def _vt_elem(vt):
    if vt >= 16:
        return "f"
    return "bfiu"[vt] if vt < 4 else "bfiu"[(vt - 4) // 3]


def _vt_count(vt):
    if vt >= 16:
        return _MAT_DIMS[vt][0] * _MAT_DIMS[vt][1]
    return 1 if vt < 4 else 2 + (vt - 4) % 3


def _matrix_array(value, a, n, as_shader_text):
    """A matrix value as an (n, C, R) array of columns.

    A literal is the C*R components in storage order = column by column (MatrixValue::to_wgsl_string, reference
    src/graph/mod.rs:1428-1441). A PROPERTY is uploaded as MatrixValue::as_bytes (graph/mod.rs:1387-1391): the first
    C*AlignOf(vecR)/4 floats of that PACKED storage, which the shader then reads as array<vecR, C> — with a 16-byte
    column stride when R = 3. For three-row matrices the two disagree, so the shader sees columns (s0 s1 s2),
    (s4 s5 s6), (s8 s9 s10) ... of the packed storage s; that is the reference's behaviour and what is restated here."""
    c, r = _MAT_DIMS[value.vt]
    if as_shader_text:
        m = a.reshape(c, r)
    else:
        stride = 2 if r == 2 else 4
        s = np.zeros(16, dtype=F32)
        s[:c * r] = a
        m = s[:c * stride].reshape(c, stride)[:, :r]
    return np.broadcast_to(m[None, :, :], (n, c, r)).copy()


def literal_array(value, n, as_shader_text=True):
    """graph.Value -> broadcast array. `as_shader_text`: the value is a literal embedded in shader code (properties
    are uploaded as bytes and keep their exact bits)."""
    elem, cnt = _vt_elem(value.vt), _vt_count(value.vt)
    w = np.array(value.words, dtype=U32)
    if elem == "f" and not as_shader_text:
        a = w.view(F32)
    elif elem == "f":
        # A float literal reaches the shader as TEXT with six decimals (ToWgslString for f32, src/lib.rs:264-269):
        # the value the reference's GPU sees is the literal rounded to 1e-6, not the literal's own bits.
        a = np.array([F32(float("%.6f" % float(x))) if np.isfinite(x) else x for x in w.view(F32)], dtype=F32)
    elif elem == "i":
        a = w.view(np.int32)
    elif elem == "b":
        a = w != 0
    else:
        a = w
    if value.vt >= 16:
        return _matrix_array(value, a, n, as_shader_text)
    if cnt == 1:
        return np.broadcast_to(a[0], (n,)).copy()
    return np.broadcast_to(a[None, :], (n, cnt)).copy()


# ---------------------------------------------------------------------------------------------------
# Expression interpreter
# ---------------------------------------------------------------------------------------------------
class Writer:
    """Counterpart of ShaderWriter for evaluation: holds the side-effect cache of one code-generation scope."""

    def __init__(self):
        self.cache = {}


class Env:
    """Everything generated code can reach: the particle record (dict name -> array), sim params, rng, ..."""

    def __init__(self, n, particle, sim, rng, props=None, particle_index=None, particle_counter=None, transform=None, parent=None,
                 parent_particle_index=None):
        self.n = n
        self.particle = particle
        self.sim = sim
        self.rng = rng
        self.props = props or {}
        self.particle_index = particle_index
        self.particle_counter = particle_counter
        self.transform = transform  # 3x4 row-major affine (rows x,y,z)
        self.parent = parent
        self.parent_particle_index = parent_particle_index
        self.is_alive = np.ones(n, dtype=bool)
        self.was_alive = np.ones(n, dtype=bool)


def ev(module, h: int, env: Env, wr: Writer):
    """Expr::eval semantics on arrays (see module docstring for the caching rule)."""
    if h in wr.cache:
        return wr.cache[h]
    node = module.nodes[h - 1]
    k = node.kind
    n = env.n
    if k == "lit":
        return literal_array(node.value, n)
    if k == "attr":
        if node.attr.name == "id":
            return env.particle_index.astype(U32)
        if node.attr.name == "particle_counter":
            return env.particle_counter.astype(U32)
        return env.particle[node.attr.name]
    if k == "parent_attr":
        if node.attr.name == "id":
            return env.parent_particle_index.astype(U32)
        return env.parent[node.attr.name]
    if k == "prop":
        return literal_array(env.props[node.prop], n, as_shader_text=False)
    if k == "builtin":
        if node.op == "rand":
            if _vt_elem(node.vt) != "f":
                raise ValueError("rand() only exists for float types")
            v = env.rng.frand_n(_vt_count(node.vt))
            wr.cache[h] = v
            return v
        if node.op == "is_alive":
            return env.is_alive
        return np.broadcast_to(F32(getattr(env.sim, node.op)), (n,)).copy()
    if k == "unary":
        a = ev(module, node.args[0], env, wr)
        return _unary(node.op, a)
    if k == "binary":
        op = node.op
        a = ev(module, node.args[0], env, wr)
        b = ev(module, node.args[1], env, wr)
        if op in ("uniform", "normal"):
            cnt = 1 if a.ndim == 1 else a.shape[1]
            if op == "uniform":
                v = a + env.rng.frand_n(cnt) * (b - a)
            else:
                u = env.rng.frand()
                vv = env.rng.frand_n(cnt)
                r = np.sqrt(F32(-2.0) * np.log(u))
                c = np.cos(TAU * vv)
                v = a + b * (r if cnt == 1 else r[:, None]) * c
            wr.cache[h] = v
            return v
        return _binary(op, a, b)
    if k == "ternary":
        a = ev(module, node.args[0], env, wr)
        b = ev(module, node.args[1], env, wr)
        c = ev(module, node.args[2], env, wr)
        if node.op == "mix":
            return mix(a, b, c)
        if node.op == "clamp":
            return clamp(a, b, c)
        if node.op == "smoothstep":
            return smoothstep(a, b, c)
        return np.stack([a, b, c], axis=1)
    if k == "cast":
        a = ev(module, node.args[0], env, wr)
        dt = _DT[_vt_elem(node.vt)]
        cnt = _vt_count(node.vt)
        a = a.astype(dt)
        if cnt > 1 and a.ndim == 1:
            a = np.repeat(a[:, None], cnt, axis=1)
        return a
    raise ValueError(k)


def _unary(op, a):
    f = {"abs": np.abs, "acos": np.arccos, "asin": np.arcsin, "atan": np.arctan, "ceil": np.ceil, "cos": np.cos, "exp": np.exp,
         "exp2": np.exp2, "floor": np.floor, "log": np.log, "log2": np.log2, "round": np.rint, "sin": np.sin, "sqrt": np.sqrt, "tan": np.tan}
    if op in f:
        return f[op](a)
    if op == "fract":
        return a - np.floor(a)
    if op == "inverse_sqrt":
        return F32(1.0) / np.sqrt(a)
    if op == "length":
        return length(a)
    if op == "normalize":
        return normalize(a)
    if op == "saturate":
        return clamp(a, F32(0.0), F32(1.0))
    if op == "sign":
        return sign(a)
    if op == "all":
        return a if a.ndim == 1 else np.all(a, axis=1)
    if op == "any":
        return a if a.ndim == 1 else np.any(a, axis=1)
    if op in "xyzw":
        return a[:, "xyzw".index(op)]
    if op == "pack4x8unorm":
        q = np.floor(F32(0.5) + F32(255.0) * clamp(a, F32(0.0), F32(1.0))).astype(U32) & U32(0xFF)
        return q[:, 0] | (q[:, 1] << U32(8)) | (q[:, 2] << U32(16)) | (q[:, 3] << U32(24))
    if op == "pack4x8snorm":
        q = np.floor(F32(0.5) + F32(127.0) * clamp(a, F32(-1.0), F32(1.0))).astype(np.int32).astype(U32) & U32(0xFF)
        return q[:, 0] | (q[:, 1] << U32(8)) | (q[:, 2] << U32(16)) | (q[:, 3] << U32(24))
    if op == "unpack4x8unorm":
        return np.stack([((a >> U32(8 * i)) & U32(0xFF)).astype(F32) / F32(255.0) for i in range(4)], axis=1)
    if op == "unpack4x8snorm":
        return np.stack([np.maximum(((a >> U32(8 * i)) & U32(0xFF)).astype(np.uint8).view(np.int8).astype(F32) / F32(127.0), F32(-1.0)) for i in range(4)], axis=1)
    raise ValueError(op)


def _matrix_mul(a, b):
    """WGSL `*` with a matrix operand ((n, C, R) arrays of columns); sums run over the columns from left to right."""
    if a.ndim == 3 and b.ndim == 1:
        return a * b[:, None, None]
    if a.ndim == 1 and b.ndim == 3:
        return a[:, None, None] * b
    if a.ndim == 3 and b.ndim == 2:                      # matCxR * vecC -> vecR
        acc = a[:, 0, :] * b[:, 0:1]
        for j in range(1, a.shape[1]):
            acc = acc + a[:, j, :] * b[:, j:j + 1]
        return acc
    if a.ndim == 2 and b.ndim == 3:                      # vecR * matCxR -> vecC
        return np.stack([dot(a, b[:, j, :]) for j in range(b.shape[1])], axis=1)
    return np.stack([_matrix_mul(a, b[:, j, :]) for j in range(b.shape[1])], axis=1)   # matKxR * matCxK -> matCxR


def _binary(op, a, b):
    if a.ndim == 3 or b.ndim == 3:
        if op == "mul":
            return _matrix_mul(a, b)
        if op in ("add", "sub") and a.shape == b.shape:
            return a + b if op == "add" else a - b
        raise ValueError(f"operator '{op}' is not defined for matrices")
    if op in ("add", "sub", "mul", "div", "rem", "gt", "ge", "lt", "le", "max", "min", "atan2", "step"):
        a, b = _bc(a, b)
    if op == "add":
        return a + b
    if op == "sub":
        return a - b
    if op == "mul":
        return a * b
    if op == "div":
        return a / b if a.dtype == F32 else a // b
    if op == "rem":
        return _trunc_rem(a, b)
    if op == "gt":
        return a > b
    if op == "ge":
        return a >= b
    if op == "lt":
        return a < b
    if op == "le":
        return a <= b
    if op == "max":
        return np.maximum(a, b)
    if op == "min":
        return np.minimum(a, b)
    if op == "atan2":
        return np.arctan2(a, b)
    if op == "step":
        return np.where(a <= b, F32(1.0), F32(0.0)).astype(F32)
    if op == "dot":
        return dot(a, b)
    if op == "cross":
        return cross(a, b)
    if op == "distance":
        return length(a - b)
    if op == "vec2":
        return np.stack([a, b], axis=1)
    if op == "vec4_xyz_w":
        return np.concatenate([a, b[:, None]], axis=1)
    raise ValueError(op)


# ---------------------------------------------------------------------------------------------------
# Modifiers (file:line in the module docstring)
# ---------------------------------------------------------------------------------------------------
def _transform_dir(env, v3):
    """(transform * vec4(v, 0)).xyz with transform = transpose(mat4x4(row0,row1,row2,(0,0,0,1))): the columns
    are summed left to right like WGSL's mat*vec."""
    t = env.transform  # (3,4) rows
    out = []
    for r in range(3):
        acc = F32(t[r, 0]) * v3[:, 0]
        acc = acc + F32(t[r, 1]) * v3[:, 1]
        acc = acc + F32(t[r, 2]) * v3[:, 2]
        acc = acc + F32(t[r, 3]) * F32(0.0)
        out.append(acc)
    return np.stack(out, axis=1)


def apply_modifier(mod, module, env: Env, wr: Writer):
    P = env.particle
    dt = F32(env.sim.delta_time)
    k = mod.kind
    E = mod.exprs

    def e(i, w=wr):
        return ev(module, E[i], env, w)

    if k == "accel":
        a = e(0)
        P["velocity"] = P["velocity"] + a * dt
    elif k == "radial_accel":
        w = Writer()
        origin = e(0, w)
        accel = e(1, w)
        radial = normalize(P["position"] - origin)
        s = accel * dt
        P["velocity"] = P["velocity"] + radial * (s[:, None] if s.ndim == 1 else s)
    elif k == "tangent_accel":
        origin, axis, accel = e(0), e(1), e(2)
        radial = normalize(P["position"] - origin)
        tangent = normalize(cross(axis, radial))
        s = accel * dt
        P["velocity"] = P["velocity"] + tangent * (s[:, None] if s.ndim == 1 else s)
    elif k == "conform_to_sphere":
        w = Writer()
        c = e(0, w)
        r = e(1, w)
        influence = e(2, w)
        shell_half = ev(module, E[5], env, w) if E[5] else np.broadcast_to(F32(0.1), (env.n,))
        max_speed = e(4, w)
        accel = e(3, w)
        sticky = ev(module, E[6], env, w) if E[6] else np.broadcast_to(F32(2.0), (env.n,))
        rel = c - P["position"]
        od = length(rel)
        direction = normalize(rel)
        sd = od - r
        active = ~(sd > influence)
        cur = dot(P["velocity"], direction)
        shell = smoothstep(F32(0.0), shell_half, np.abs(sd))
        maxr = sign(sd) * shell * max_speed
        ds = maxr - cur
        sticky_accel = accel * sticky
        acc = mix(sticky_accel, accel, shell)
        cds = dt * acc
        imp = sign(ds) * np.minimum(np.abs(ds), cds)
        newv = P["velocity"] + imp[:, None] * direction
        P["velocity"] = np.where(active[:, None], newv, P["velocity"])
    elif k == "linear_drag":
        drag = e(0)
        f = np.maximum(F32(0.0), F32(1.0) - (drag * dt))
        P["velocity"] = P["velocity"] * (f[:, None] if f.ndim == 1 else f)
    elif k == "kill_sphere":
        diff = P["position"] - e(0)
        sq = dot(diff, diff)
        r2 = e(1)
        cond = (sq < r2) if mod.params[0] else (sq > r2)
        env.is_alive = np.where(cond, False, env.is_alive)
    elif k == "kill_aabb":
        dist = np.abs(P["position"] - e(0))
        half = e(1)
        cond = np.all(dist < half, axis=1) if mod.params[0] else np.any(dist > half, axis=1)
        env.is_alive = np.where(cond, False, env.is_alive)
    elif k == "set_attribute":
        from bevy_hanabi_b200.graph import ATTRIBUTES
        P[ATTRIBUTES[mod.params[0]].name] = np.array(e(0), copy=True)
    elif k == "inherit_attribute":
        from bevy_hanabi_b200.graph import ATTRIBUTES
        name = ATTRIBUTES[mod.params[0]].name
        P[name] = np.array(env.parent[name], copy=True)
    elif k == "set_position_circle":
        w = Writer()
        c = e(0, w)
        nrm = e(1, w)
        rad = e(2, w)  # operand draws are hoisted before the body's own frand() calls
        r = np.sqrt(env.rng.frand()) * rad if mod.params[0] == 1 else rad  # volume: radius drawn BEFORE theta
        sg = step(np.broadcast_to(F32(0.0), (env.n,)), nrm[:, 2]) * F32(2.0) - F32(1.0)
        a = F32(-1.0) / (sg + nrm[:, 2])
        b = nrm[:, 0] * nrm[:, 1] * a
        tangent = np.stack([F32(1.0) + sg * nrm[:, 0] * nrm[:, 0] * a, sg * b, -sg * nrm[:, 0]], axis=1)
        bitangent = np.stack([b, sg + nrm[:, 1] * nrm[:, 1] * a, -nrm[:, 1]], axis=1)
        theta = env.rng.frand() * TAU
        direction = tangent * np.cos(theta)[:, None] + bitangent * np.sin(theta)[:, None]
        P["position"] = c + r[:, None] * direction
    elif k == "set_position_sphere":
        w = Writer()
        c = e(0, w)
        rad = e(1, w)
        r = np.power(env.rng.frand(), F32(1.0) / F32(3.0)) * rad if mod.params[0] == 1 else rad
        theta = env.rng.frand() * TAU
        z = env.rng.frand() * F32(2.0) - F32(1.0)
        phi = np.arccos(z)
        sinphi = np.sin(phi)
        x = sinphi * np.cos(theta)
        y = sinphi * np.sin(theta)
        direction = np.stack([x, y, z], axis=1)
        P["position"] = c + r[:, None] * direction
    elif k == "set_position_cone3d":
        w = Writer()
        h0 = e(0, w)
        rt = e(2, w)
        rb = e(1, w)
        alpha_h = np.power(env.rng.frand(), F32(1.0) / F32(3.0))
        h = h0 * alpha_h
        r0 = rb + (rt - rb) * alpha_h
        alpha_r = np.sqrt(env.rng.frand())
        r = r0 * alpha_r
        theta = env.rng.frand() * TAU
        p = np.stack([r * np.cos(theta), h, r * np.sin(theta)], axis=1)
        P["position"] = _transform_dir(env, p)
    elif k == "set_velocity_circle":
        w = Writer()
        c = e(0, w)
        axis = e(1, w)
        speed = e(2, w)
        delta = P["position"] - c
        radial = normalize(delta - dot(delta, axis)[:, None] * axis)
        v = _transform_dir(env, radial)
        P["velocity"] = v * (speed[:, None] if speed.ndim == 1 else speed)
    elif k == "set_velocity_sphere":
        c = e(0)
        speed = e(1)
        P["velocity"] = normalize(P["position"] - c) * (speed[:, None] if speed.ndim == 1 else speed)
    elif k == "set_velocity_tangent":
        w = Writer()
        o = e(0, w)
        axis = e(1, w)
        speed = e(2, w)
        radial = P["position"] - o
        tangent = normalize(cross(axis, radial))
        v = _transform_dir(env, tangent)
        P["velocity"] = v * (speed[:, None] if speed.ndim == 1 else speed)
    elif k == "emit_spawn_event":
        count = e(0).astype(U32)
        cond = env.is_alive if mod.params[0] == 0 else (env.was_alive & ~env.is_alive)
        env.emitted.append((mod.params[1], np.where(cond, count, U32(0))))
    else:
        raise ValueError(k)


# ---------------------------------------------------------------------------------------------------
# Effect-level oracle operating on tests.helpers.RefWorld (reference layouts)
# ---------------------------------------------------------------------------------------------------
class EffectOracle:
    def __init__(self, asset, properties: dict | None = None):
        self.asset = asset
        self.module = asset.module
        fields, self.stride, _ = asset.particle_layout()
        self.fields = [f for f in fields if not f.name.startswith("pad")]
        self.names = {f.name for f in self.fields}
        self.stride_words = self.stride // 4
        # per-instance property values (dict name -> graph.Value); defaults from the module
        self.default_props = {name: v for name, v in self.module.properties}
        self.props = properties

    # -- AoS <-> dict
    def unpack(self, rec: np.ndarray) -> dict:
        out = {}
        for f in self.fields:
            cnt = _vt_count(f.vt)
            w = rec[:, f.offset // 4: f.offset // 4 + cnt]
            elem = _vt_elem(f.vt)
            a = w.view(F32) if elem == "f" else (w.view(np.int32) if elem == "i" else w)
            out[f.name] = np.ascontiguousarray(a[:, 0] if cnt == 1 else a)
        return out

    def pack(self, P: dict, rec: np.ndarray, skip=()):
        for f in self.fields:
            if f.name in skip:
                continue
            cnt = _vt_count(f.vt)
            a = np.asarray(P[f.name])
            a = a.reshape(len(rec), cnt)
            rec[:, f.offset // 4: f.offset // 4 + cnt] = np.ascontiguousarray(a).view(U32)

    def _props_for(self, world, inst_index):
        vals = dict(self.default_props)
        if self.props and inst_index in self.props:
            from bevy_hanabi_b200.graph import Value
            for k, v in self.props[inst_index].items():
                vals[k] = Value.of(v)
        return vals

    # -- passes
    def init_pass(self, world, b: int = 0):
        members = world.batches[b]
        total = world.batch_spawn_total(b)
        if total == 0:
            return
        threads = (total + 63) // 64 * 64
        cpu_prefix = []
        run = 0
        for i in members:
            cpu_prefix.append(run)
            run += max(0, world.spawners[i].spawn)
        for j, i in enumerate(members):
            sp, md = world.spawners[i], world.metadata[i]
            end = cpu_prefix[j + 1] if j + 1 < len(members) else threads
            rng_threads = max(0, end - cpu_prefix[j])
            n = min(rng_threads, sp.spawn & 0xFFFFFFFF, md.max_spawn)
            if n == 0:
                continue
            base = sp.slab_offset
            k = np.arange(n, dtype=np.int64)
            alive_index = md.alive_count + k
            slots = world.indirect[base + alive_index, 2].astype(np.int64)
            pidx = (slots - base).astype(U32)
            counter = (md.particle_counter + k).astype(U32)
            rng = Rng(pcg_hash(pidx ^ U32(sp.seed)))
            rec = np.zeros((n, self.stride_words), dtype=U32)
            P = self.unpack(rec)
            tr = np.array(list(sp.transform), dtype=F32).reshape(3, 4)
            env = Env(n, P, world.sim, rng, self._props_for(world, i), pidx, counter, tr)
            wr = Writer()
            for m in self.asset.init_modifiers:
                apply_modifier(m, self.module, env, wr)
            if "prev" in self.names:
                P["prev"] = np.full(n, 0xFFFFFFFF, dtype=U32)
            if "next" in self.names:
                P["next"] = np.full(n, 0xFFFFFFFF, dtype=U32)
            if self.asset.simulation_space == 0:  # Global: particle.position += transform[3].xyz (lib.rs:525-528)
                P["position"] = P["position"] + tr[:, 3][None, :]
            self.pack(P, rec)
            W = md.indirect_write_index
            world.indirect[base + alive_index, W] = pidx
            world.particles[base + pidx.astype(np.int64)] = rec
            md.alive_count += n
            md.particle_counter = (md.particle_counter + n) & 0xFFFFFFFF

    def update_pass(self, world, b: int = 0):
        self.last_emitted = []  # [(channel, per-row event counts)] of the last instance updated (none when nothing was alive)
        for i in world.batches[b]:
            sp, md = world.spawners[i], world.metadata[i]
            n = md.max_update
            if n == 0:
                continue
            base = sp.slab_offset
            W = md.indirect_write_index
            R = 1 - W
            pidx = world.indirect[base:base + n, R].copy()
            if getattr(world, "slot_order", False):
                # HNB_EFFECT_SLOT_ORDER: the threads visit the particles in ascending particle index — the reference's update
                # (vfx_update.wgsl:106-167) when its alive list happens to be sorted; the list itself is left as it is
                pidx = np.sort(pidx)
            rows = base + pidx.astype(np.int64)
            rec = world.particles[rows].copy()
            P = self.unpack(rec)
            rng = Rng(pcg_hash(pidx ^ U32(sp.seed)))
            tr = np.array(list(sp.transform), dtype=F32).reshape(3, 4)
            env = Env(n, P, world.sim, rng, self._props_for(world, i), pidx, np.zeros(n, dtype=U32), tr)
            env.emitted = []
            dt = F32(world.sim.delta_time)
            # AGE_CODE / REAP_CODE (lib.rs:1223-1264)
            if "age" in self.names:
                if "lifetime" in self.names:
                    env.was_alive = P["age"] < P["lifetime"]
                P["age"] = P["age"] + dt
                if "lifetime" in self.names:
                    env.is_alive = P["age"] < P["lifetime"]
                    env.is_alive = env.is_alive & (P["age"] < P["lifetime"])
            motion = self.asset.motion_integration
            euler = motion != 0 and "position" in self.names and "velocity" in self.names
            if euler and motion == 1:
                P["position"] = P["position"] + P["velocity"] * dt
            wr = Writer()
            for m in self.asset.update_modifiers:
                apply_modifier(m, self.module, env, wr)
            if euler and motion == 2:
                P["position"] = P["position"] + P["velocity"] * dt
            self.pack(P, rec, skip=("prev", "next"))
            world.particles[rows] = rec
            alive = env.is_alive
            n_alive = int(alive.sum())
            n_dead = n - n_alive
            inst = int(world.draw[5 * md.indirect_render_index + 1])
            world.indirect[base + inst: base + inst + n_alive, W] = pidx[alive]
            world.draw[5 * md.indirect_render_index + 1] = inst + n_alive
            if n_dead:
                kk = np.arange(n_dead, dtype=np.int64)
                alive_index = md.alive_count - 1 - kk
                world.indirect[base + alive_index, 2] = (base + pidx[~alive].astype(np.int64)).astype(U32)
                md.alive_count -= n_dead
                md.max_spawn += n_dead
            self.last_emitted = env.emitted

    def frame(self, world, orc_c):
        """init -> indirect -> prefix sum -> update [-> ribbon sort], bookkeeping passes by the C oracle."""
        for b in range(len(world.batches)):
            self.init_pass(world, b)
        world.oracle_indirect(orc_c)
        world.oracle_prefix_sum(orc_c)
        for b in range(len(world.batches)):
            self.update_pass(world, b)
        if "ribbon_id" in self.names:  # LayoutFlags::RIBBONS (lib.rs:1018-1019) -> sort passes (mod.rs:7372-7610)
            world.oracle_sort_ribbons(orc_c)
