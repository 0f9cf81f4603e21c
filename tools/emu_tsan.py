"""ThreadSanitizer over the emulated kernels (tests/kernel_emu.py, tests/static_emu.py): stand-alone C++ programs run the
ribbon-sort kernels and the real
hnb_update text of the C5 effect for a few frames on 2 CTAs of OS threads, built with -fsanitize=thread. TSAN knows
pthread barriers (= __syncwarp / __syncthreads / the collectives) and __atomic operations (= the tile states, tickets),
so any report is a plain memory access pair of the KERNEL that is not ordered by them — the CPU analogue of
compute-sanitizer racecheck, but for global memory and the shared-memory stash alike.

    python tools/emu_tsan.py            # prints the TSAN summary; exit code 0 when clean
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bevy_hanabi_b200 import recipes  # noqa: E402
from tests import kernel_emu as K  # noqa: E402

MAIN = r"""
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char** argv) {
    const uint32_t rows = 6000, chunks = argc > 1 ? atoi(argv[1]) : 2, ctas = argc > 2 ? atoi(argv[2]) : 2, frames = 4;
    const uint32_t tile = 32u * HNB_TILE_K * chunks;
    const uint32_t tile_word = tile, small = tile;
    std::vector<float4> plane0(rows), plane1(rows);
    std::vector<uint32_t> ping(rows), pong(rows), dead(rows), tile_prefix(2), prefix_sum(1), spawn_prefix(1), batch_tiles(1), ticket(1), draw(5);
    std::vector<unsigned long long> states(rows / small + 4);
    uint32_t s = 12345u;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return float(s >> 8) / 16777216.0f; };
    const uint32_t alive0 = 5600;
    for (uint32_t i = 0; i < rows; ++i) {
        plane0[i] = make_float4(rnd(), rnd(), rnd(), 0.f);
        plane1[i] = make_float4(rnd() - .5f, rnd() - .5f, rnd() - .5f, 0.01f + 0.08f * rnd());  // lifetimes: many deaths per frame
        ping[i] = pong[i] = i < alive0 ? i : 0u;
        dead[i] = i;
    }
    hnb::FrameHeader frame; memset(&frame, 0, sizeof frame);
    frame.sim.delta_time = 1.f / 60.f; frame.sim.num_effects = 1; frame.num_batches = 1;
    hnb::Spawner sp; memset(&sp, 0, sizeof sp);
    sp.transform[0] = sp.transform[5] = sp.transform[10] = 1.f; sp.inverse_transform[0] = sp.inverse_transform[5] = sp.inverse_transform[10] = 1.f;
    sp.seed = 42; sp.parent_slab_offset = 0xFFFFFFFFu;
    hnb::EffectMetadata md; memset(&md, 0xFF, sizeof md);
    md.capacity = rows; md.alive_count = alive0; md.max_update = 0; md.max_spawn = rows - alive0; md.indirect_write_index = 0; md.indirect_render_index = 0;
    md.particle_stride = 8; md.particle_counter = 0;
    hnb::BatchInfo bi = {0, 0, 0, 0, 0, 1};
    EmuBatch b; memset(&b, 0, sizeof b);
    b.frame = &frame; b.spawners = &sp; b.spawn_prefix = spawn_prefix.data(); b.prefix_sum = prefix_sum.data(); b.tile_prefix = tile_prefix.data();
    b.batch_info = &bi; b.batch_tiles = batch_tiles.data(); b.ticket = ticket.data(); b.tile_state = states.data(); b.metadata = &md; b.draw_args = draw.data();
    b.planes[0] = plane0.data(); b.planes[1] = plane1.data(); b.ping = ping.data(); b.pong = pong.data(); b.dead = dead.data();
    b.capacity = rows; b.tile_rows = tile_word;
    for (uint32_t f = 0; f < frames; ++f) {
        frame.epoch = f + 1; frame.sim.time = f * frame.sim.delta_time;
        // vfx_indirect + vfx_prefix_sum for one instance (restated; the emulation covers the per-particle kernels)
        draw[1] = 0; md.max_update = md.alive_count; md.max_spawn = md.capacity - md.alive_count;
        md.indirect_write_index = 1u - md.indirect_write_index; sp.render_indirect_read_index = md.indirect_write_index;
        prefix_sum[0] = 0; bi.total_update_count = md.alive_count;
        const uint32_t tiles = hnb::hnb_tile_count(md.alive_count, tile_word);
        tile_prefix[0] = 0; tile_prefix[1] = tiles; batch_tiles[0] = tiles; ticket[0] = 0;
        emu_update(&b, ctas, 64 * 1024);
        printf("frame %u: updated %u -> alive %u (instance_count %u)\n", f, md.max_update, md.alive_count, draw[1]);
        if (md.alive_count != draw[1]) return 2;
    }
    return 0;
}
"""


SORT_MAIN = r"""
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
int main(int argc, char** argv) {
    // one slab, two instances: 5000 keys (cooperative radix sort) and 900 keys (shared-memory bitonic sort)
    const uint32_t counts[2] = {5000, 900}, rows = 6000, grid = argc > 1 ? atoi(argv[1]) : 2;
    const bool wide = argc > 2 && atoi(argv[2]);
    std::vector<uint32_t> plane(rows * 4), ping(rows), pong(rows);
    uint32_t s = 99u;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return s; };
    for (uint32_t i = 0; i < rows; ++i) {
        plane[4 * i + 0] = wide ? rnd() : rnd() % 5u;                          // ribbon id (word 0)
        const float age = float(rnd() >> 8) / 16777216.0f;
        plane[4 * i + 1] = wide ? rnd() : __float_as_uint(age);                 // age (word 1)
    }
    hnb::Spawner sp[2]; hnb::EffectMetadata md[2];
    memset(sp, 0, sizeof sp); memset(md, 0, sizeof md);
    uint32_t off = 0;
    for (int i = 0; i < 2; ++i) {
        for (uint32_t r = 0; r < counts[i]; ++r) ping[off + r] = (r * 7919u) % counts[i];   // a permutation (7919 is prime)
        md[i].capacity = counts[i]; md[i].alive_count = counts[i]; md[i].indirect_write_index = 0; md[i].sort_key_offset = 0; md[i].sort_key2_offset = 1;
        sp[i].effect_metadata_index = i; sp[i].slab_offset = off;
        off += counts[i];
    }
    hnb::RibbonSortArgs a; memset(&a, 0, sizeof a);
    a.planes.ptr[0] = plane.data(); a.planes.words[0] = 4; a.planes.word_off[0] = 0;
    for (int w = 0; w < 4; ++w) a.planes.word_to_plane[w] = 0;
    a.ping = ping.data(); a.pong = pong.data(); a.spawners = sp; a.metadata = md; a.spawner_base = 0; a.instance_count = 2;
    std::vector<unsigned long long> k0(rows), k1(rows); std::vector<uint32_t> v0(rows), v1(rows), hist(2 * 8 * 256 + 256 * grid);
    a.scratch_keys[0] = k0.data(); a.scratch_keys[1] = k1.data(); a.scratch_vals[0] = v0.data(); a.scratch_vals[1] = v1.data();
    a.scratch_hist = hist.data(); a.scratch_rows = rows; a.scratch_grid = grid;
    semu_ribbon_sort_small(&a);
    semu_ribbon_sort_large(&a, grid);
    off = 0;
    for (int i = 0; i < 2; ++i) {
        unsigned long long prev = 0;
        for (uint32_t r = 0; r < counts[i]; ++r) {
            const uint32_t row = off + ping[off + r];
            const unsigned long long key = ((unsigned long long)plane[4 * row] << 32) | plane[4 * row + 1];
            if (key < prev) { printf("instance %d not sorted at %u\n", i, r); return 2; }
            prev = key;
        }
        printf("instance %d: %u keys sorted\n", i, counts[i]);
        off += counts[i];
    }
    return 0;
}
"""


EVENTS_MAIN = r"""
#include <stdio.h>
#include <stdlib.h>
int main() {
    const uint32_t rows = 9000, capacity_rows = 9500, cap = 4000;
    std::vector<uint32_t> counts(capacity_rows), ping(capacity_rows), pong(capacity_rows), buffer(cap, 0xDEADBEEFu), block_sums(capacity_rows / 2048 + 2);
    uint32_t s = 7u, total = 0;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return s; };
    for (uint32_t i = 0; i < capacity_rows; ++i) { counts[i] = (rnd() >> 8) % 3u; pong[i] = i * 3u; if (i < rows) total += counts[i]; }
    hnb::EffectMetadata md; memset(&md, 0, sizeof md); md.max_update = rows; md.indirect_write_index = 0; md.base_child_index = 0;
    hnb::Spawner sp; memset(&sp, 0, sizeof sp);
    hnb::ChildInfo ci[1] = {{0, 0}};
    hnb::EventAppendArgs a; memset(&a, 0, sizeof a);
    a.counts = counts.data(); a.ping = ping.data(); a.pong = pong.data(); a.spawner = &sp; a.metadata = &md; a.block_sums = block_sums.data();
    a.child_infos = ci; a.binding = 0; a.buffer = buffer.data(); a.capacity = cap;
    semu_ordered_event_append(&a, capacity_rows);
    uint32_t pos = 0;
    for (uint32_t r = 0; r < rows && pos < cap; ++r)
        for (uint32_t i = 0; i < counts[r] && pos < cap; ++i, ++pos)
            if (buffer[pos] != r * 3u) { printf("event %u wrong\n", pos); return 2; }
    printf("ordered append: %u events requested, %d counted, %u kept and in order\n", total, ci[0].event_count, pos);
    return ci[0].event_count == (int)total ? 0 : 3;
}
"""


def static_program(name, main_text):
    """Build the emulated static kernels + `main_text` with -fsanitize=thread."""
    from tests import static_emu as S
    import re
    wgsl = S._strip_includes((S.KERNELS / "hnb_wgsl.cuh").read_text())
    tables = S._strip_includes((S.KERNELS / "hnb_tables.cuh").read_text())
    header = S._strip_includes((S.KERNELS / "hnb_static_kernels.h").read_text())
    static = (S.KERNELS / "hnb_static_kernels.cu").read_text()
    static = S._strip_includes(static[:static.index("cudaError_t launch_indirect(")]) + "\n}  // namespace hnb\n"
    ribbon = (S.KERNELS / "hnb_ribbon_sort.cu").read_text()
    ribbon = S._strip_includes(ribbon[:ribbon.index("cudaError_t launch_ribbon_sort(")]) + "\n}  // namespace hnb\n"
    static = re.sub(r"__global__ void k_measure_sm_clock.*?\n}\n", "", static, flags=re.S)
    body = S._rewrite_shared(static + "\n" + ribbon)
    out = ROOT / "build" / "kernel_emu"
    out.mkdir(parents=True, exist_ok=True)
    cpp, exe = out / f"tsan_{name}.cpp", out / f"tsan_{name}"
    cpp.write_text(K.PRELUDE + S.EXTRA_PRELUDE + wgsl + "\n" + tables + "\n" + header + "\n" + body + S.DRIVER + main_text)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-w", str(cpp), "-o", str(exe)], check=True)
    return exe


def events_program():
    exe = static_program("events", EVENTS_MAIN)
    p = subprocess.run([str(exe)], capture_output=True, text=True, env={"TSAN_OPTIONS": "halt_on_error=0 report_signal_unsafe=0"})
    races = p.stderr.count("WARNING: ThreadSanitizer: data race")
    print(f"ordered event append kernels: exit {p.returncode}, {races} data-race reports")
    print(p.stdout.strip())
    if races:
        print(p.stderr[:6000])
    return p.returncode or races


def sort_program():
    """The ribbon-sort kernels (bitonic in shared memory + cooperative radix sort) under ThreadSanitizer."""
    from tests import static_emu as S
    wgsl = S._strip_includes((S.KERNELS / "hnb_wgsl.cuh").read_text())
    tables = S._strip_includes((S.KERNELS / "hnb_tables.cuh").read_text())
    header = S._strip_includes((S.KERNELS / "hnb_static_kernels.h").read_text())
    static = (S.KERNELS / "hnb_static_kernels.cu").read_text()
    static = S._strip_includes(static[:static.index("cudaError_t launch_indirect(")]) + "\n}  // namespace hnb\n"
    ribbon = (S.KERNELS / "hnb_ribbon_sort.cu").read_text()
    ribbon = S._strip_includes(ribbon[:ribbon.index("cudaError_t launch_ribbon_sort(")]) + "\n}  // namespace hnb\n"
    import re
    static = re.sub(r"__global__ void k_measure_sm_clock.*?\n}\n", "", static, flags=re.S)
    body = S._rewrite_shared(static + "\n" + ribbon)
    out = ROOT / "build" / "kernel_emu"
    out.mkdir(parents=True, exist_ok=True)
    cpp, exe = out / "tsan_sort.cpp", out / "tsan_sort"
    cpp.write_text(K.PRELUDE + S.EXTRA_PRELUDE + wgsl + "\n" + tables + "\n" + header + "\n" + body + S.DRIVER + SORT_MAIN)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-w", str(cpp), "-o", str(exe)], check=True)
    rc = 0
    for grid, wide in ((2, 0), (3, 1)):
        p = subprocess.run([str(exe), str(grid), str(wide)], capture_output=True, text=True, env={"TSAN_OPTIONS": "halt_on_error=0 report_signal_unsafe=0"})
        races = p.stderr.count("WARNING: ThreadSanitizer: data race")
        print(f"ribbon sort grid={grid} wide_keys={wide}: exit {p.returncode}, {races} data-race reports")
        print(p.stdout.strip())
        if races:
            print(p.stderr[:6000])
        rc |= p.returncode or races
    return rc


def main():
    rc_sort = sort_program() | events_program()
    src = recipes.c5_lowered().generate_source()
    for old, new in K.SUBSTITUTIONS:
        assert src.count(old) == 1
        src = src.replace(old, new)
    out = ROOT / "build" / "kernel_emu"
    out.mkdir(parents=True, exist_ok=True)
    cpp, exe = out / "tsan_c5.cpp", out / "tsan_c5"
    cpp.write_text(K.PRELUDE + src + K.DRIVER + MAIN)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-pthread", "-w", str(cpp), "-o", str(exe)], check=True)
    rc = 0
    for chunks, ctas in ((1, 2), (2, 2), (4, 1)):
        p = subprocess.run([str(exe), str(chunks), str(ctas)], capture_output=True, text=True, env={"TSAN_OPTIONS": "halt_on_error=0 report_signal_unsafe=0"})
        races = p.stderr.count("WARNING: ThreadSanitizer: data race")
        print(f"chunks={chunks} ctas={ctas}: exit {p.returncode}, {races} data-race reports")
        print(p.stdout.strip())
        if races:
            print(p.stderr[:6000])
        rc |= p.returncode or races
    return 1 if (rc or rc_sort) else 0


if __name__ == "__main__":
    sys.exit(main())
