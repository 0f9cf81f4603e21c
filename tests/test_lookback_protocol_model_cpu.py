"""A small executable model of hnb_update's scheduling protocol (bevy_hanabi_b200/csrc/kernels/hnb_particle_kernels.cuh):

  * a persistent grid whose CTAs may be only PARTLY resident (other batches of the same frame run concurrently on
    side streams; a CTA that is not resident holds no ticket);
  * in-order tile tickets: one request per CTA for its warps' first tiles, one per warp afterwards, requested
    after the tile's pass 1;
  * a state word per tile: AGGREGATE(count) published right after pass 1 — PREFIX(count) for the first tile of an
    instance — and PREFIX(inclusive prefix) once resolved;
  * the look-back: windows of 32 predecessors; a window is consumed when every entry up to its first PREFIX (or
    all 32) is published, tiles before the instance's first tile count as PREFIX(0);
  * deferred compaction: a warp resolves tile t only after streaming its next tile (or when it runs out of tickets).

The model runs the warps as coroutines under a randomised scheduler and checks, for many shapes and schedules, that
the protocol always terminates (no warp ever waits on something only a non-resident or a waiting warp could
provide) and that every tile gets exactly the exclusive prefix of its instance. It documents WHY the kernel is
deadlock-free; it does not execute the CUDA code (the GPU suite does that).
"""
import random

import pytest

WINDOW = 32
pytestmark = pytest.mark.timeout(600)  # a protocol bug must fail, not hang


class Grid:
    def __init__(self, alive_per_tile, inst_first, warps_per_cta, n_ctas, resident_ctas, rng, defer=True):
        self.alive = alive_per_tile
        self.inst_first = inst_first          # per tile: first tile of its instance
        self.total = len(alive_per_tile)
        self.state = [None] * self.total      # None | ("A", v) | ("P", v)
        self.ticket = 0
        self.exclusive = [None] * self.total
        self.rng = rng
        self.defer = defer
        self.W = warps_per_cta
        self.waiting_ctas = list(range(n_ctas))
        self.resident_limit = resident_ctas
        self.running = {}                     # cta -> list of live warp coroutines
        self.polls = 0

    # ---- one warp --------------------------------------------------------------------------------------
    def warp(self, first_tile):
        tile, pending = first_tile, None
        while tile < self.total:
            yield "pass1"                                             # stream the rows (no dependence on anyone)
            a = self.alive[tile]
            self.state[tile] = ("P", a) if tile == self.inst_first[tile] else ("A", a)
            yield "published"
            nxt = self.ticket                                         # atomicAdd(ticket, 1)
            self.ticket += 1
            if self.defer:
                if pending is not None:
                    yield from self.resolve(pending)
                pending = tile
            else:
                yield from self.resolve(tile)
            tile = nxt
        if pending is not None:
            yield from self.resolve(pending)

    def resolve(self, t):
        first = self.inst_first[t]
        if t == first:
            self.exclusive[t] = 0
            return
        total, pos = 0, t - 1
        while True:
            window = []
            for back in range(WINDOW):
                window.append(self.state[pos - back] if pos - back >= first else ("P", 0))
            first_p = next((i for i, s in enumerate(window) if s is not None and s[0] == "P"), None)
            need = window[:first_p + 1] if first_p is not None else window
            if any(s is None for s in need):
                self.polls += 1
                yield "poll"                                          # spin: some needed predecessor has not published
                continue
            total += sum(s[1] for s in need)
            if first_p is not None:
                break
            pos -= WINDOW
        self.exclusive[t] = total
        self.state[t] = ("P", total + self.alive[t])
        yield "resolved"

    # ---- the machine -----------------------------------------------------------------------------------
    def admit(self):
        while self.waiting_ctas and len(self.running) < self.resident_limit:
            cta = self.waiting_ctas.pop(0)                            # the hardware starts CTAs in any order it likes;
            base = self.ticket                                        # whichever starts takes the NEXT tickets
            self.ticket += self.W
            self.running[cta] = [self.warp(base + w) for w in range(self.W)]

    def run(self, max_steps=2_000_000):
        self.rng.shuffle(self.waiting_ctas)
        self.admit()
        steps, idle = 0, 0
        while self.running:
            cta = self.rng.choice(list(self.running))
            warps = self.running[cta]
            w = self.rng.randrange(len(warps))
            try:
                ev = next(warps[w])
                idle = idle + 1 if ev == "poll" else 0
            except StopIteration:
                warps.pop(w)
                idle = 0
                if not warps:
                    del self.running[cta]
                    self.admit()
            steps += 1
            assert steps < max_steps, "protocol did not terminate"
            # if every resident warp only polls for this long, nobody can make progress any more: deadlock
            assert idle < 50_000, "deadlock: all resident warps are waiting"
        assert not self.waiting_ctas


def _instances_to_tiles(rng, n_instances, max_tiles_per_instance, tile_rows=128):
    alive, inst_first = [], []
    for _ in range(n_instances):
        tiles = rng.randint(0, max_tiles_per_instance)
        first = len(alive)
        for _ in range(tiles):
            alive.append(rng.randint(0, tile_rows))
            inst_first.append(first)
    return alive, inst_first


def _check(grid):
    run = 0
    for t in range(grid.total):
        if grid.inst_first[t] == t:
            run = 0
        assert grid.exclusive[t] == run, f"tile {t}: exclusive prefix {grid.exclusive[t]} != {run}"
        run += grid.alive[t]


@pytest.mark.parametrize("seed", range(12))
def test_protocol_terminates_with_exact_prefixes(seed):
    rng = random.Random(seed)
    n_instances = rng.choice([1, 1, 2, 5, 40])
    alive, inst_first = _instances_to_tiles(rng, n_instances, rng.choice([3, 40, 150]))
    warps_per_cta = rng.choice([1, 2, 8])
    n_ctas = rng.choice([1, 3, 16])
    resident = rng.randint(1, n_ctas)          # partial residency, down to a single CTA
    g = Grid(alive, inst_first, warps_per_cta, n_ctas, resident, rng, defer=True)
    g.run()
    _check(g)


@pytest.mark.parametrize("defer", [True, False])
def test_long_chain_single_resident_cta(defer):
    """One instance of 300 tiles (the look-back crosses many 32-tile windows), 16 CTAs of which only ONE is resident
    at a time: the in-order tickets are what makes this terminate."""
    rng = random.Random(99)
    alive = [rng.randint(0, 128) for _ in range(300)]
    g = Grid(alive, [0] * 300, 8, 16, 1, rng, defer=defer)
    g.run()
    _check(g)


def test_static_tile_assignment_would_deadlock():
    """Counter-example motivating the ticket: if tile t were bound to grid warp t up front (no ticket), a resident CTA
    can wait for a tile owned by a CTA that cannot become resident until the first one exits."""
    class StaticGrid(Grid):
        def admit(self):
            while self.waiting_ctas and len(self.running) < self.resident_limit:
                cta = self.waiting_ctas.pop(0)
                self.running[cta] = [self.warp_static(cta * self.W + w) for w in range(self.W)]

        def warp_static(self, tile):
            if tile < self.total:
                yield "pass1"
                a = self.alive[tile]
                self.state[tile] = ("P", a) if tile == self.inst_first[tile] else ("A", a)
                yield from self.resolve(tile)

    rng = random.Random(1)
    g = StaticGrid([5] * 16, [0] * 16, 2, 8, 1, rng)
    g.waiting_ctas = [3, 0, 1, 2, 4, 5, 6, 7]   # the scheduler happens to start CTA 3 first
    g.rng = random.Random(2)
    g.rng.shuffle = lambda x: None               # keep that order
    with pytest.raises(AssertionError, match="deadlock"):
        g.run()
