/*
 * host_producers_c_api.c — the host-side producers of the per-frame tables from plain C, no GPU needed:
 *   node graph (reference src/graph/node.rs)   -> one expression -> a SetAttribute update modifier -> generated code
 *   EffectProperties store (src/properties.rs) -> the Properties blob hnb_upload_properties takes
 *   EffectSimulation clock (src/time.rs)       -> the GpuSimParams record hnb_set_sim_params takes
 *
 *   gcc -O2 -Iinclude examples/host_producers_c_api.c -Lbevy_hanabi_b200 -lhanabi_b200 -Wl,-rpath,$PWD/bevy_hanabi_b200 -o build/host_producers_c_api
 */
#include <stdio.h>
#include <string.h>

#include "hanabi_b200.h"
#include "hanabi_b200_graph.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int32_t rc_ = (call);                                                \
        if (rc_ != HNB_OK) {                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, hnb_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

static uint32_t bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static uint32_t attribute_index(const char* name) {
    for (uint32_t i = 0; i < hnb_attribute_count(); ++i) {
        const char* n;
        uint32_t vt, def[4];
        hnb_attribute_info(i, &n, &vt, def);
        if (strcmp(n, name) == 0) return i;
    }
    return 0xFFFFFFFFu;
}

int main(void) {
    /* ---- node graph: position + velocity * delta_time (the wiring of node.rs `graph`, :944-968) */
    hnb_node_graph* g = hnb_node_graph_create();
    uint32_t n_pos = hnb_node_graph_add_node(g, HNB_NODE_ATTRIBUTE, "position");
    uint32_t n_add = hnb_node_graph_add_node(g, HNB_NODE_ADD, NULL);
    uint32_t n_vel = hnb_node_graph_add_node(g, HNB_NODE_ATTRIBUTE, "velocity");
    uint32_t n_mul = hnb_node_graph_add_node(g, HNB_NODE_MUL, NULL);
    uint32_t n_time = hnb_node_graph_add_node(g, HNB_NODE_TIME, NULL);
    uint32_t add_in[2], mul_in[2], n;
    CHECK(hnb_node_graph_slots(g, n_add, 1, add_in, 2, &n));
    CHECK(hnb_node_graph_slots(g, n_mul, 1, mul_in, 2, &n));
    CHECK(hnb_node_graph_link(g, hnb_node_graph_find_slot(g, n_pos, 2, "position"), add_in[0]));
    CHECK(hnb_node_graph_link(g, hnb_node_graph_find_slot(g, n_vel, 2, "velocity"), mul_in[0]));
    CHECK(hnb_node_graph_link(g, hnb_node_graph_find_slot(g, n_time, 2, "delta_time"), mul_in[1]));
    CHECK(hnb_node_graph_link(g, hnb_node_graph_find_slot(g, n_mul, 2, "result"), add_in[1]));

    hnb_module* m = hnb_module_create();
    uint32_t speed_default[1] = {bits(2.0f)}, zero3[3] = {0, 0, 0}, up3[3] = {0, bits(1.0f), 0};
    hnb_prop p_speed = hnb_module_add_property(m, "speed", HNB_FLOAT, speed_default);
    hnb_expr euler = 0;
    CHECK(hnb_node_graph_eval_slot(g, m, hnb_node_graph_find_slot(g, n_add, 2, "result"), &euler));
    char text[512], stmts[512];
    CHECK(hnb_module_eval(m, euler, 2, text, sizeof text, stmts, sizeof stmts));
    printf("graph lowers to: %s\n", text);
    hnb_expr_info info;
    CHECK(hnb_module_get(m, euler, &info));
    printf("root expression: kind %u op %u operands %u %u (of %u)\n", info.kind, info.op, info.operands[0], info.operands[1], hnb_module_len(m));

    /* ---- an asset using it: init position / velocity, update position through the graph (no built-in integration) */
    hnb_expr zero = hnb_module_lit(m, HNB_VEC3, zero3);
    hnb_expr vel = hnb_module_binary(m, HNB_BIN_MUL, hnb_module_lit(m, HNB_VEC3, up3), hnb_module_prop(m, p_speed));
    hnb_asset* asset = hnb_asset_create("graph_authored", 1024, m);
    uint32_t a_pos = attribute_index("position"), a_vel = attribute_index("velocity");
    CHECK(hnb_asset_set_motion_integration(asset, 0));
    CHECK(hnb_asset_add_modifier(asset, 1, HNB_MOD_SET_ATTRIBUTE, &zero, 1, &a_pos, 1));
    CHECK(hnb_asset_add_modifier(asset, 1, HNB_MOD_SET_ATTRIBUTE, &vel, 1, &a_vel, 1));
    CHECK(hnb_asset_add_modifier(asset, 2, HNB_MOD_SET_ATTRIBUTE, &euler, 1, &a_pos, 1));
    hnb_generated* gen = NULL;
    CHECK(hnb_asset_generate(asset, NULL, 0, &gen));
    hnb_effect_desc desc;
    CHECK(hnb_generated_desc(gen, &desc));
    printf("update code: %s\n", desc.update_code);

    /* ---- EffectProperties: stored values win over the asset's defaults, unknown names are dropped by update() */
    hnb_effect_properties* props = hnb_effect_properties_create();
    uint32_t w_speed[1] = {bits(7.5f)}, w_junk[1] = {bits(1.0f)}, changed = 0, blob_size = 0;
    CHECK(hnb_effect_properties_set(props, "speed", HNB_FLOAT, w_speed));
    CHECK(hnb_effect_properties_set(props, "not_in_the_asset", HNB_FLOAT, w_junk));
    CHECK(hnb_effect_properties_update(props, asset, &changed));
    float blob[4] = {0};
    CHECK(hnb_effect_properties_serialize(props, asset, blob, sizeof blob, &blob_size));
    printf("properties: %u stored, changed %u, blob %u bytes, speed = %g\n", hnb_effect_properties_len(props), changed, blob_size, blob[0]);
    if (hnb_effect_properties_set(props, "speed", HNB_VEC3, zero3) == HNB_OK) return 2; /* type mismatch must be refused */
    printf("type mismatch refused: %s\n", hnb_last_error());

    /* ---- Time<EffectSimulation>: 60 Hz frames at half speed, one paused */
    hnb_sim_clock* clock = hnb_sim_clock_create();
    CHECK(hnb_sim_clock_set_relative_speed(clock, 0.5));
    hnb_sim_params sim;
    for (int f = 0; f < 3; ++f) {
        if (f == 2) hnb_sim_clock_pause(clock);
        CHECK(hnb_sim_clock_advance(clock, 16666667ull));
        CHECK(hnb_sim_clock_sim_params(clock, 1, &sim));
        printf("frame %d: delta_time %.9g time %.9g real_time %.9g was_paused %u\n", f, sim.delta_time, sim.time, sim.real_time, hnb_sim_clock_was_paused(clock));
    }
    if (hnb_sim_clock_set_relative_speed(clock, -1.0) == HNB_OK) return 2;
    printf("negative speed refused: %s\n", hnb_last_error());

    hnb_sim_clock_destroy(clock);
    hnb_effect_properties_destroy(props);
    hnb_generated_destroy(gen);
    hnb_asset_destroy(asset);
    hnb_module_destroy(m);
    hnb_node_graph_destroy(g);
    return 0;
}
