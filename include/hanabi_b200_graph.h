/*
 * hanabi_b200_graph.h — Level-2 C ABI: effect authoring and lowering to CUDA C.
 *
 * Replaces, for the simulation passes, the reference's src/graph (Module / Expr), src/modifier,
 * src/attributes.rs (ParticleLayout), src/properties.rs (PropertyLayout) and
 * EffectShaderSources::generate (src/lib.rs:805-1335). It exists in this library because the Rust
 * toolchain is absent from the build image; the result of hnb_asset_generate() is an hnb_effect_desc
 * that Level 1 (hanabi_b200.h, hnb_effect_compile) consumes. Pure CPU: no function here needs a GPU.
 *
 * Expression and property handles are 1-based; 0 means "invalid" and is returned on error together
 * with a message in hnb_last_error().
 */
#ifndef HANABI_B200_GRAPH_H
#define HANABI_B200_GRAPH_H

#include "hanabi_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hnb_module hnb_module; /* Module, reference src/graph/expr.rs:337 */
typedef struct hnb_asset hnb_asset;   /* EffectAsset, reference src/asset.rs:272 */
typedef uint32_t hnb_expr;
typedef uint32_t hnb_prop;

/* BuiltInOperator (expr.rs:1580) */
enum { HNB_BUILTIN_TIME = 0, HNB_BUILTIN_DELTA_TIME, HNB_BUILTIN_VIRTUAL_TIME, HNB_BUILTIN_VIRTUAL_DELTA_TIME, HNB_BUILTIN_REAL_TIME,
       HNB_BUILTIN_REAL_DELTA_TIME, HNB_BUILTIN_RAND, HNB_BUILTIN_ALPHA_CUTOFF, HNB_BUILTIN_IS_ALIVE };
/* UnaryOperator (expr.rs:1833), same order */
enum { HNB_UN_ABS = 0, HNB_UN_ACOS, HNB_UN_ASIN, HNB_UN_ATAN, HNB_UN_ALL, HNB_UN_ANY, HNB_UN_CEIL, HNB_UN_COS, HNB_UN_EXP, HNB_UN_EXP2,
       HNB_UN_FLOOR, HNB_UN_FRACT, HNB_UN_INV_SQRT, HNB_UN_LENGTH, HNB_UN_LOG, HNB_UN_LOG2, HNB_UN_NORMALIZE, HNB_UN_PACK4X8SNORM,
       HNB_UN_PACK4X8UNORM, HNB_UN_ROUND, HNB_UN_SATURATE, HNB_UN_SIGN, HNB_UN_SIN, HNB_UN_SQRT, HNB_UN_TAN, HNB_UN_UNPACK4X8SNORM,
       HNB_UN_UNPACK4X8UNORM, HNB_UN_W, HNB_UN_X, HNB_UN_Y, HNB_UN_Z };
/* BinaryOperator (expr.rs:2079), same order */
enum { HNB_BIN_ADD = 0, HNB_BIN_ATAN2, HNB_BIN_CROSS, HNB_BIN_DISTANCE, HNB_BIN_DIV, HNB_BIN_DOT, HNB_BIN_GT, HNB_BIN_GE, HNB_BIN_LT,
       HNB_BIN_LE, HNB_BIN_MAX, HNB_BIN_MIN, HNB_BIN_MUL, HNB_BIN_REM, HNB_BIN_STEP, HNB_BIN_SUB, HNB_BIN_UNIFORM_RAND,
       HNB_BIN_NORMAL_RAND, HNB_BIN_VEC2, HNB_BIN_VEC4_XYZ_W };
/* TernaryOperator (expr.rs:2306) */
enum { HNB_TER_MIX = 0, HNB_TER_CLAMP, HNB_TER_SMOOTHSTEP, HNB_TER_VEC3 };

/* ---- Module --------------------------------------------------------------------------------- */
HNB_API hnb_module* hnb_module_create(void);
HNB_API void hnb_module_destroy(hnb_module* m);
/** Literal of `value_type` (hnb_value_type); `words` holds its 32-bit lanes (bools: 0 / non-zero; a matCxR: C*R floats,
 *  column by column, like MatrixValue::new in the reference's src/graph/mod.rs:1283-1311). */
HNB_API hnb_expr hnb_module_lit(hnb_module* m, uint32_t value_type, const uint32_t* words);
/** One stored expression, as Module::get returns it (expr.rs:607-612). `kind`: 0 built-in, 1 literal, 2 property,
 *  3 attribute, 4 parent attribute, 5 unary, 6 binary, 7 ternary, 8 cast (the order of the reference's Expr enum,
 *  expr.rs:910-995, without TextureSample). `op`: the HNB_BUILTIN_* / HNB_UN_* / HNB_BIN_* / HNB_TER_* code. */
typedef struct hnb_expr_info {
    uint32_t kind, op;
    uint32_t value_type;        /* literal type, rand type, cast target */
    hnb_expr operands[3];
    hnb_prop property;
    const char* attribute;      /* static attribute name */
    uint32_t literal_words[16];
} hnb_expr_info;
HNB_API uint32_t hnb_module_len(const hnb_module* m);
HNB_API int32_t hnb_module_get(const hnb_module* m, hnb_expr e, hnb_expr_info* out);
HNB_API hnb_expr hnb_module_attr(hnb_module* m, const char* attribute_name);
HNB_API hnb_expr hnb_module_parent_attr(hnb_module* m, const char* attribute_name);
/** A matrix property larger than 16 bytes must end up as the last entry of the property layout (the reference's
 *  PropertyLayout::new, properties.rs:561-699, cannot place anything after it); hnb_asset_generate reports a violation. */
HNB_API hnb_prop hnb_module_add_property(hnb_module* m, const char* name, uint32_t value_type, const uint32_t* default_words);
HNB_API hnb_expr hnb_module_prop(hnb_module* m, hnb_prop property);
/** `rand_value_type` is only read for HNB_BUILTIN_RAND. */
HNB_API hnb_expr hnb_module_builtin(hnb_module* m, uint32_t op, uint32_t rand_value_type);
HNB_API hnb_expr hnb_module_unary(hnb_module* m, uint32_t op, hnb_expr e);
HNB_API hnb_expr hnb_module_binary(hnb_module* m, uint32_t op, hnb_expr left, hnb_expr right);
HNB_API hnb_expr hnb_module_ternary(hnb_module* m, uint32_t op, hnb_expr a, hnb_expr b, hnb_expr c);
HNB_API hnb_expr hnb_module_cast(hnb_module* m, hnb_expr e, uint32_t target_value_type);
HNB_API int32_t hnb_module_is_const(const hnb_module* m, hnb_expr e);
HNB_API int32_t hnb_module_has_side_effect(const hnb_module* m, hnb_expr e);
/**
 * Expr::eval in a fresh ShaderWriter of the given context (1 = init, 2 = update) with the module's
 * own property layout and a layout made of every attribute: writes the expression text to `out`
 * and the hoisted side-effect statements to `stmts` (either may be NULL). ≙ the reference's
 * expression-text unit tests (src/graph/expr.rs:4219-4300).
 */
HNB_API int32_t hnb_module_eval(const hnb_module* m, hnb_expr e, uint32_t context, char* out, size_t out_cap, char* stmts, size_t stmts_cap);

/* ---- Attributes / layouts ------------------------------------------------------------------- */
HNB_API uint32_t hnb_attribute_count(void);
/** Name / value type / default value lanes of built-in attribute `index` (reference attributes.rs:1338-1378 order). */
HNB_API int32_t hnb_attribute_info(uint32_t index, const char** name, uint32_t* value_type, uint32_t default_words[4]);
/** ParticleLayoutBuilder::build for a set of attribute names. `out` receives up to `cap` fields in offset order
 *  (padding fields are named pad0..pad4 and included); *n = field count, *size = record bytes, *align. */
HNB_API int32_t hnb_particle_layout_build(const char* const* attribute_names, uint32_t n_names, hnb_attr_layout* out, uint32_t cap,
                                          uint32_t* n, uint32_t* size, uint32_t* align);
/** f32 literal formatting (ToWgslString for f32, reference src/lib.rs:264-269, plus the C suffix). */
HNB_API int32_t hnb_format_f32(float value, char* out, size_t cap);

/* ---- Asset ---------------------------------------------------------------------------------- */
enum { HNB_CONTEXT_INIT = 1, HNB_CONTEXT_UPDATE = 2 };
/* Modifier kinds; operand order = field order of the reference struct (optional operands may be 0). */
enum {
    HNB_MOD_ACCEL = 1,             /* exprs: accel */
    HNB_MOD_RADIAL_ACCEL,          /* origin, accel */
    HNB_MOD_TANGENT_ACCEL,         /* origin, axis, accel */
    HNB_MOD_CONFORM_TO_SPHERE,     /* origin, radius, influence_dist, attraction_accel, max_attraction_speed, [shell_half_thickness], [sticky_factor] */
    HNB_MOD_LINEAR_DRAG,           /* drag */
    HNB_MOD_KILL_SPHERE,           /* center, sqr_radius; params: kill_inside */
    HNB_MOD_KILL_AABB,             /* center, half_size; params: kill_inside */
    HNB_MOD_SET_ATTRIBUTE,         /* value; params: attribute index */
    HNB_MOD_INHERIT_ATTRIBUTE,     /* params: attribute index */
    HNB_MOD_SET_POSITION_CIRCLE,   /* center, axis, radius; params: dimension (0 surface, 1 volume) */
    HNB_MOD_SET_POSITION_SPHERE,   /* center, radius; params: dimension */
    HNB_MOD_SET_POSITION_CONE3D,   /* height, base_radius, top_radius; params: dimension */
    HNB_MOD_SET_VELOCITY_CIRCLE,   /* center, axis, speed */
    HNB_MOD_SET_VELOCITY_SPHERE,   /* center, speed */
    HNB_MOD_SET_VELOCITY_TANGENT,  /* origin, axis, speed */
    HNB_MOD_EMIT_SPAWN_EVENT       /* count; params: condition (0 always, 1 on die), child_index */
};
/** The module is copied into the asset (EffectAsset::new takes ownership of the Module). */
HNB_API hnb_asset* hnb_asset_create(const char* name, uint32_t capacity, const hnb_module* module);
HNB_API void hnb_asset_destroy(hnb_asset* a);
HNB_API int32_t hnb_asset_set_simulation_space(hnb_asset* a, uint32_t local);        /* SimulationSpace */
HNB_API int32_t hnb_asset_set_motion_integration(hnb_asset* a, uint32_t mode);       /* 0 none, 1 pre, 2 post */
HNB_API int32_t hnb_asset_add_modifier(hnb_asset* a, uint32_t context, uint32_t kind, const hnb_expr* exprs, uint32_t n_exprs,
                                       const uint32_t* params, uint32_t n_params);
/** EffectAsset::particle_layout (asset.rs:605-626); same output convention as hnb_particle_layout_build. */
HNB_API int32_t hnb_asset_particle_layout(const hnb_asset* a, hnb_attr_layout* out, uint32_t cap, uint32_t* n, uint32_t* size, uint32_t* align);
/** PropertyLayout of the asset's module: entries in layout order; *size = bytes of one Properties record. */
HNB_API int32_t hnb_asset_property_layout(const hnb_asset* a, hnb_attr_layout* out, uint32_t cap, uint32_t* n, uint32_t* size);
/** EffectProperties::serialize: `names[i]` is set to the value lanes `words[i]` (others keep their default). */
HNB_API int32_t hnb_asset_serialize_properties(const hnb_asset* a, const char* const* names, const uint32_t* const* words, uint32_t n,
                                               void* blob, uint32_t blob_cap, uint32_t* blob_size);

/* ---- Node-graph front end of the expression module (reference src/graph/node.rs) ------------------- */
enum hnb_node_kind {
    HNB_NODE_ADD = 1,   /* AddNode: lhs, rhs -> result (node.rs:467-506) */
    HNB_NODE_SUB,       /* SubNode (node.rs:509-549) */
    HNB_NODE_MUL,       /* MulNode (node.rs:552-592) */
    HNB_NODE_DIV,       /* DivNode (node.rs:595-635) */
    HNB_NODE_ATTRIBUTE, /* AttributeNode: -> <attribute name> (node.rs:638-694) */
    HNB_NODE_TIME,      /* TimeNode: -> time, delta_time (node.rs:697-733) */
    HNB_NODE_NORMALIZE  /* NormalizeNode: one input expression -> out (node.rs:736-775) */
};
typedef struct hnb_node_graph hnb_node_graph; /* Graph, node.rs:244 */
HNB_API hnb_node_graph* hnb_node_graph_create(void);
HNB_API void hnb_node_graph_destroy(hnb_node_graph* g);
/** Graph::add_node (node.rs:284-310): NodeId (1-based), 0 on error. `attribute`: HNB_NODE_ATTRIBUTE only (NULL = position). */
HNB_API uint32_t hnb_node_graph_add_node(hnb_node_graph* g, uint32_t kind, const char* attribute);
HNB_API uint32_t hnb_node_graph_node_count(const hnb_node_graph* g);
/** Graph::link / unlink / unlink_all (node.rs:313-352) on SlotIds (1-based). Where the reference asserts on the slot
 *  direction the call returns HNB_ERR_EXPR. An input slot keeps one source: linking it again replaces the source. */
HNB_API int32_t hnb_node_graph_link(hnb_node_graph* g, uint32_t output_slot, uint32_t input_slot);
HNB_API int32_t hnb_node_graph_unlink(hnb_node_graph* g, uint32_t output_slot, uint32_t input_slot);
HNB_API int32_t hnb_node_graph_unlink_all(hnb_node_graph* g, uint32_t slot);
/** Graph::slots / input_slots / output_slots (node.rs:355-420): dir 0 = all, 1 = inputs, 2 = outputs. */
HNB_API int32_t hnb_node_graph_slots(const hnb_node_graph* g, uint32_t node, uint32_t dir, uint32_t* out, uint32_t cap, uint32_t* n);
/** Graph::input_slot (dir 1) / output_slot (dir 2) by name; node 0 and dir 0: Graph::get_slot_id. 0 = none. */
HNB_API uint32_t hnb_node_graph_find_slot(const hnb_node_graph* g, uint32_t node, uint32_t dir, const char* name);
/** Slot accessors (node.rs:144-198); *value_type = -1 for an untyped slot. */
HNB_API int32_t hnb_node_graph_slot_info(const hnb_node_graph* g, uint32_t slot, const char** name, uint32_t* node, uint32_t* is_input,
                                         int32_t* value_type, uint32_t* linked, uint32_t cap, uint32_t* n_linked);
/** Node::eval (node.rs:458-463): lower one node into `m` from explicit input expressions. */
HNB_API int32_t hnb_node_graph_eval_node(const hnb_node_graph* g, uint32_t node, hnb_module* m, const hnb_expr* inputs, uint32_t n_inputs,
                                         hnb_expr* outputs, uint32_t cap, uint32_t* n_outputs);
/** Lower everything an output slot depends on (each node once, in dependency order; unlinked inputs and cycles
 *  are errors). The reference leaves this walk to the caller. */
HNB_API int32_t hnb_node_graph_eval_slot(hnb_node_graph* g, hnb_module* m, uint32_t output_slot, hnb_expr* out);

/* ---- EffectProperties: per-instance property values (reference src/properties.rs:205-454) ---------- */
typedef struct hnb_effect_properties hnb_effect_properties;
HNB_API hnb_effect_properties* hnb_effect_properties_create(void);
HNB_API void hnb_effect_properties_destroy(hnb_effect_properties* p);
HNB_API uint32_t hnb_effect_properties_len(const hnb_effect_properties* p);
/** EffectProperties::set (properties.rs:319-341): overwrite, or append a new property whose default is `value`.
 *  Where the reference asserts (value of another type than the property's) the call returns HNB_ERR_EXPR with the
 *  reference's message and changes nothing. `words`: the value's 32-bit lanes, as for hnb_module_lit. */
HNB_API int32_t hnb_effect_properties_set(hnb_effect_properties* p, const char* name, uint32_t value_type, const uint32_t* words);
/** EffectProperties::set_if_changed (properties.rs:343-376): *changed = 0 when the stored value was already equal. */
HNB_API int32_t hnb_effect_properties_set_if_changed(hnb_effect_properties* p, const char* name, uint32_t value_type,
                                                     const uint32_t* words, uint32_t* changed);
/** EffectProperties::get_stored (properties.rs:305-310): 1 and the value (16 lanes) when present, else 0. */
HNB_API int32_t hnb_effect_properties_get_stored(const hnb_effect_properties* p, const char* name, uint32_t* value_type, uint32_t* words16);
/** The i-th PropertyInstance: name (valid until the store changes), type, current value, default value. */
HNB_API int32_t hnb_effect_properties_get(const hnb_effect_properties* p, uint32_t index, const char** name, uint32_t* value_type,
                                          uint32_t* value_words16, uint32_t* default_words16);
/** EffectProperties::update (properties.rs:378-417) against the properties of the asset's module. *changed tells
 *  whether the store was mutated (what Bevy's change detection would record). */
HNB_API int32_t hnb_effect_properties_update(hnb_effect_properties* p, const hnb_asset* asset, uint32_t* changed);
/** EffectProperties::serialize (properties.rs:437-453) with the asset's PropertyLayout: the blob for
 *  hnb_upload_properties. Same output convention as hnb_asset_serialize_properties. */
HNB_API int32_t hnb_effect_properties_serialize(const hnb_effect_properties* p, const hnb_asset* asset, void* blob, uint32_t blob_cap,
                                                uint32_t* blob_size);

typedef struct hnb_generated hnb_generated; /* EffectShaderSources for the simulation passes */
/** EffectShaderSources::generate. `parent` (or NULL) provides the parent particle layout of a GPU-event child. */
HNB_API int32_t hnb_asset_generate(const hnb_asset* a, const hnb_asset* parent, uint32_t num_event_bindings, hnb_generated** out);
/** Fill `desc` with pointers into `g` (valid until hnb_generated_destroy). */
HNB_API int32_t hnb_generated_desc(const hnb_generated* g, hnb_effect_desc* desc);
HNB_API void hnb_generated_destroy(hnb_generated* g);

/* ---- CPU producers feeding the hot path ----------------------------------------- */
/** SpawnerSettings (reference src/spawn.rs:219): CpuValue<f32> fields as [lo, hi] (lo == hi: Single). */
typedef struct hnb_spawner_settings {
    float count_lo, count_hi;
    float spawn_duration_lo, spawn_duration_hi;
    float period_lo, period_hi;
    uint32_t cycle_count;   /* 0 = forever, 1 = once */
    uint32_t starts_active;
    uint32_t emit_on_start;
} hnb_spawner_settings;
typedef struct hnb_effect_spawner hnb_effect_spawner; /* EffectSpawner, spawn.rs:640 */
typedef struct hnb_effect_spawner_state_t {
    float cycle_time, cycle_spawn_duration, cycle_period, cycle_ratio, cycle_spawn_count;
    uint32_t completed_cycle_count, active, has_completed, spawn_count;
} hnb_effect_spawner_state_t;
/** SpawnerSettings::try_new (spawn.rs:313-338): validates the period like the reference. */
HNB_API int32_t hnb_spawner_settings_new(float count_lo, float count_hi, float duration_lo, float duration_hi, float period_lo,
                                         float period_hi, uint32_t cycle_count, hnb_spawner_settings* out);
HNB_API int32_t hnb_spawner_settings_once(float count, hnb_spawner_settings* out);                /* spawn.rs:349 */
HNB_API int32_t hnb_spawner_settings_rate(float rate, hnb_spawner_settings* out);                 /* spawn.rs:367 */
HNB_API int32_t hnb_spawner_settings_burst(float count, float period, hnb_spawner_settings* out); /* spawn.rs:381 */
HNB_API hnb_effect_spawner* hnb_effect_spawner_create(const hnb_spawner_settings* settings, uint64_t rng_seed);
HNB_API void hnb_effect_spawner_destroy(hnb_effect_spawner* s);
/** EffectSpawner::tick (spawn.rs:838-921): number of particles to spawn this frame -> GpuSpawnerParams.spawn. */
HNB_API int32_t hnb_effect_spawner_tick(hnb_effect_spawner* s, float dt, uint32_t* spawn_count);
HNB_API void hnb_effect_spawner_reset(hnb_effect_spawner* s);
HNB_API void hnb_effect_spawner_set_active(hnb_effect_spawner* s, uint32_t active);
HNB_API int32_t hnb_effect_spawner_state(const hnb_effect_spawner* s, hnb_effect_spawner_state_t* out);

/* ---- Simulation clock -> GpuSimParams (reference src/time.rs, src/render/mod.rs:193-279, :2796-2811) ---- */
/** Time<Real>, Time<Virtual> (bevy_time 0.19, restated) and Time<EffectSimulation> (time.rs:30-46), in integer
 *  nanoseconds like Rust's Duration. */
typedef struct hnb_sim_clock hnb_sim_clock;
typedef struct hnb_sim_clock_state_t {
    uint64_t real_elapsed_ns, real_delta_ns;
    uint64_t virtual_elapsed_ns, virtual_delta_ns;
    uint64_t sim_elapsed_ns, sim_delta_ns;
    double virtual_effective_speed; /* Time<Virtual>::effective_speed_f64 */
    double sim_effective_speed;     /* EffectSimulationTime::effective_speed_f64 (time.rs:127) */
} hnb_sim_clock_state_t;
HNB_API hnb_sim_clock* hnb_sim_clock_create(void);
HNB_API void hnb_sim_clock_destroy(hnb_sim_clock* c);
/** Time<Virtual>::set_relative_speed_f64 / pause / unpause / set_max_delta (default 250 ms). */
HNB_API int32_t hnb_sim_clock_set_virtual_relative_speed(hnb_sim_clock* c, double ratio);
HNB_API void hnb_sim_clock_set_virtual_paused(hnb_sim_clock* c, uint32_t paused);
HNB_API int32_t hnb_sim_clock_set_max_delta_ns(hnb_sim_clock* c, uint64_t ns);
/** EffectSimulationTime (time.rs:48-162). Where the reference asserts (non-finite or negative ratio,
 *  time.rs:138-139) the call returns HNB_ERR_INVALID_ARG with the same message and changes nothing. */
HNB_API int32_t hnb_sim_clock_set_relative_speed(hnb_sim_clock* c, double ratio);
HNB_API void hnb_sim_clock_pause(hnb_sim_clock* c);
HNB_API void hnb_sim_clock_unpause(hnb_sim_clock* c);
HNB_API uint32_t hnb_sim_clock_is_paused(const hnb_sim_clock* c);
HNB_API uint32_t hnb_sim_clock_was_paused(const hnb_sim_clock* c);
HNB_API double hnb_sim_clock_relative_speed(const hnb_sim_clock* c);
HNB_API double hnb_sim_clock_effective_speed(const hnb_sim_clock* c);
/** One frame: time_system (Real, Virtual) then effect_simulation_time_system (time.rs:164-183), given the real
 *  time that passed since the previous frame. */
HNB_API int32_t hnb_sim_clock_advance(hnb_sim_clock* c, uint64_t real_delta_ns);
/** extract_sim_params (mod.rs:2796-2811) + From<&SimParams> for GpuSimParams (mod.rs:266-279): the record to hand
 *  to hnb_set_sim_params. */
HNB_API int32_t hnb_sim_clock_sim_params(const hnb_sim_clock* c, uint32_t num_effects, hnb_sim_params* out);
HNB_API void hnb_sim_params_default(hnb_sim_params* out); /* GpuSimParams::default, mod.rs:244-256 */
HNB_API int32_t hnb_sim_clock_state(const hnb_sim_clock* c, hnb_sim_clock_state_t* out);

/** What EffectBatch::try_merge compares (reference src/render/batch.rs:153-173). */
typedef struct hnb_batch_key {
    uint64_t asset_id;       /* handle */
    uint32_t slab_id;
    uint32_t pipeline_id;    /* init_and_update_pipeline_ids */
    uint32_t property_key;
    uint32_t parent_slab_id; /* 0xFFFFFFFF if none */
    uint32_t uses_gpu_events;/* cached_effect_events.is_some() */
    uint32_t is_cpu_spawner; /* spawn_info.is_cpu() */
} hnb_batch_key;
typedef struct hnb_batcher hnb_batcher; /* Batcher, batch.rs:197 */
HNB_API hnb_batcher* hnb_batcher_create(void);
HNB_API void hnb_batcher_destroy(hnb_batcher* b);
HNB_API void hnb_batcher_clear(hnb_batcher* b);
/** Batcher::push (batch.rs:348-386): *batch_index = index of the new batch, or -1 if merged into the last one. */
HNB_API int32_t hnb_batcher_push(hnb_batcher* b, const hnb_batch_key* key, uint32_t spawner_base, uint32_t slab_offset,
                                 uint32_t instance_spawn_count, int32_t* batch_index);
/** Close the open batch and expose the tables to upload with hnb_upload_batches (pointers valid until the next
 *  push/clear) plus each batch's CPU total spawn count (for hnb_batch_launch.total_spawn_count). */
HNB_API int32_t hnb_batcher_finish(hnb_batcher* b, const hnb_batch_info** infos, uint32_t* n_batches, const uint32_t** prefix,
                                   uint32_t* n_prefix, uint32_t* total_spawn_counts, uint32_t total_cap);

/** EffectSorter (reference src/render/batch.rs:476-637): the order in which batch_effects() visits the instances —
 *  dependency level (children before their parents, pinned by batch.rs:776-826), then slab, then row offset. */
#define HNB_NO_ENTITY 0xFFFFFFFFFFFFFFFFull
typedef struct hnb_effect_sorter hnb_effect_sorter;
HNB_API hnb_effect_sorter* hnb_effect_sorter_create(void);
HNB_API void hnb_effect_sorter_destroy(hnb_effect_sorter* s);
HNB_API void hnb_effect_sorter_insert(hnb_effect_sorter* s, uint64_t entity, uint32_t slab_id, uint32_t base_instance,
                                      uint64_t parent /* HNB_NO_ENTITY if none */);
HNB_API int32_t hnb_effect_sorter_sort(hnb_effect_sorter* s); /* -1: unknown parent or cycle */
HNB_API uint32_t hnb_effect_sorter_len(const hnb_effect_sorter* s);
HNB_API uint64_t hnb_effect_sorter_get(const hnb_effect_sorter* s, uint32_t index);

/* ------------------------------------------------------------------------------------ */
/* Placement of effect instances into slabs ≙ ParticleSlab / EffectCache bookkeeping      */
/* (reference src/render/effect_cache.rs:484-607, :843-930). Host only: the caller owns  */
/* the device storage (hnb_slab_create / _destroy / _reset_rows in hanabi_b200.h).        */
/* ------------------------------------------------------------------------------------ */
#define HNB_SLAB_MIN_CAPACITY 65536u /* ParticleSlab::MIN_CAPACITY, effect_cache.rs:234 */
enum { HNB_SLAB_USED = 0, HNB_SLAB_FREE = 1 }; /* SlabState */

/** Slice allocator of one slab: ParticleSlab::{allocate, free_slice} (effect_cache.rs:540-607). */
typedef struct hnb_slice_allocator hnb_slice_allocator;
HNB_API hnb_slice_allocator* hnb_slice_allocator_create(uint32_t capacity); /* capacity = max(capacity, 65536) */
HNB_API void hnb_slice_allocator_destroy(hnb_slice_allocator* s);
HNB_API uint32_t hnb_slice_allocator_capacity(const hnb_slice_allocator* s);
HNB_API uint32_t hnb_slice_allocator_used_size(const hnb_slice_allocator* s);
HNB_API uint32_t hnb_slice_allocator_free_count(const hnb_slice_allocator* s);
HNB_API int32_t hnb_slice_allocator_free_range(const hnb_slice_allocator* s, uint32_t index, uint32_t* start, uint32_t* end);
/** 0 and [*start, *end) on success, -1 when the slab has no room (allocate() -> None). */
HNB_API int32_t hnb_slice_allocator_allocate(hnb_slice_allocator* s, uint32_t size, uint32_t* start, uint32_t* end);
/** HNB_SLAB_FREE when this was the last allocated slice, else HNB_SLAB_USED. */
HNB_API int32_t hnb_slice_allocator_free(hnb_slice_allocator* s, uint32_t start, uint32_t end);

/** CachedEffect (effect_cache.rs:624-632) plus what the caller needs to create the storage. */
typedef struct hnb_cached_effect {
    uint32_t slab_index;     /* SlabId */
    uint32_t range_start;    /* slice.range: rows of the slab = spawner.slab_offset .. */
    uint32_t range_end;
    uint32_t slab_capacity;  /* rows of the slab (for hnb_slab_create when `created`) */
    uint32_t created;        /* 1: a new slab was opened for this instance */
} hnb_cached_effect;
typedef struct hnb_effect_cache hnb_effect_cache;
HNB_API hnb_effect_cache* hnb_effect_cache_create(void);
HNB_API void hnb_effect_cache_destroy(hnb_effect_cache* c);
HNB_API uint32_t hnb_effect_cache_slab_count(const hnb_effect_cache* c);   /* slots, live or not */
HNB_API int32_t hnb_effect_cache_slab_is_live(const hnb_effect_cache* c, uint32_t slab_index);
/** EffectCache::insert (effect_cache.rs:843-914). */
HNB_API int32_t hnb_effect_cache_insert(hnb_effect_cache* c, uint64_t asset_id, uint32_t capacity, hnb_cached_effect* out);
/** EffectCache::remove (:918-938): HNB_SLAB_FREE when the slab became empty (destroy its storage), -1 on a bad handle. */
HNB_API int32_t hnb_effect_cache_remove(hnb_effect_cache* c, const hnb_cached_effect* effect);

#ifdef __cplusplus
}
#endif
#endif /* HANABI_B200_GRAPH_H */
