"""(CPU only) What a tile taper can buy, from the measured timeline of `hnb_update` — a model, NOT a measurement.

Dynamic tickets hand tiles to 3552 resident warps; a warp streams a tile in a time proportional to its rows plus a fixed
per-tile cost, and — per `tools/diag_timeline.py` (profiles/r1_perf_matrix.txt): warps of an 8 Mi-particle launch end over a
26 us window out of 113 us, i.e. one 512-row tile time — a warp's tile time does not shrink when other warps have run
dry (latency-bound per warp). The kernel ends with the last warp. The model replays that schedule with and without
`HNB_TILE_TAPER` and prints the predicted kernel time; the per-tile cost is the free parameter (0.3 / 1 / 2 us).

    python tools/taper_model.py
"""
import heapq

WARPS = 3552
ROWS_US = 512 / 23.0     # rows per microsecond of one warp under load: 8 Mi rows / 3552 warps in ~106 us of streaming
RAMP_US = 8.0            # launch ramp before the first pass 1 completes work at full rate (timeline: 3.5 us start + first loads)


def makespan(rows, tiles, overhead_us):
    """tiles: list of row counts in ticket order."""
    heap = [RAMP_US] * WARPS
    heapq.heapify(heap)
    for r in tiles:
        t = heapq.heappop(heap)
        heapq.heappush(heap, t + overhead_us + r / ROWS_US)
    return max(heap)


def tiles_of(rows, S, taper_tiles=0, shift=0):
    if not taper_tiles:
        n = -(-rows // S)
        return [S] * (n - 1) + [rows - (n - 1) * S]
    n_big = (rows - min(rows, taper_tiles * S)) // S
    s = S >> shift
    rem = rows - n_big * S
    n_small = -(-rem // s)
    return [S] * n_big + [s] * (n_small - 1) + [rem - (n_small - 1) * s]


def main():
    print("rows      tile  | plain      | taper 100 % -> 128 rows | taper 50 % | taper 200 %   (us, per-tile cost 0.3 / 1 / 2 us)")
    for mi in (1, 2, 4, 8, 16, 64):
        rows = mi << 20
        S = 512 if rows >= 4 * WARPS * 128 else 256
        shift = 2 if S == 512 else 1
        cols = []
        for pct in (0, 100, 50, 200):
            cols.append(" / ".join(f"{makespan(rows, tiles_of(rows, S, WARPS * pct // 100, shift), o):6.1f}" for o in (0.3, 1.0, 2.0)))
        ideal = RAMP_US + rows / ROWS_US / WARPS
        print(f"{mi:3d} Mi  {S:4d}  | " + " | ".join(cols) + f"   (perfectly balanced: {ideal:6.1f})")


if __name__ == "__main__":
    main()
