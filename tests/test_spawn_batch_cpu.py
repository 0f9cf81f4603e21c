"""CPU producers feeding the hot path: EffectSpawner::tick against the reference's own test sequences
(src/spawn.rs:1046-1287) and Batcher::push / try_merge (src/render/batch.rs:153-188, :265-386)."""
import pytest

from bevy_hanabi_b200._native import HanabiError
from bevy_hanabi_b200.spawn import BatchKey, Batcher, EffectSpawner, SpawnerSettings


def test_new_sequence():
    # spawn.rs test_new: count 3 over 3 s, period 10 s, 2 cycles
    sp = EffectSpawner(SpawnerSettings(3.0, 3.0, 10.0, 2))
    assert sp.tick(2.0) == 2
    s = sp.state
    assert (s.cycle_time, s.cycle_spawn_duration, s.cycle_period, s.cycle_spawn_count, s.completed_cycle_count) == (2.0, 3.0, 10.0, 3.0, 0)
    assert s.cycle_ratio == pytest.approx(0.2, abs=0) or abs(s.cycle_ratio - 0.2) < 1e-7
    assert sp.tick(5.0) == 1
    assert sp.state.cycle_time == 7.0
    assert sp.tick(8.0) == 3
    s = sp.state
    assert (s.cycle_time, s.completed_cycle_count) == (5.0, 1)
    assert sp.tick(10.0) == 0
    assert sp.state.completed_cycle_count == 2 and sp.active
    assert sp.tick(0.1) == 0
    assert sp.state.completed_cycle_count == 2


def test_period_validation():
    with pytest.raises(HanabiError):
        SpawnerSettings(3.0, 1.0, (-1.0, 1.0), 0)       # negative period
    with pytest.raises(HanabiError):
        SpawnerSettings(3.0, 1.0, (0.0, 0.0), 0)        # cannot generate a positive period
    with pytest.raises(HanabiError):
        SpawnerSettings(3.0, 1.0, (0.0, float("inf")), 0)
    SpawnerSettings(3.0, 1.0, (0.0, 0.0), 1)            # once(): period unchecked


def test_once():
    sp = EffectSpawner(SpawnerSettings.once(5.0))
    assert sp.active
    assert sp.tick(0.001) == 5
    assert sp.tick(100.0) == 0
    sp = EffectSpawner(SpawnerSettings.once(5.0))
    sp.tick(1.0)
    sp.reset()
    assert sp.tick(1.0) == 5


def test_once_start_inactive():
    sp = EffectSpawner(SpawnerSettings.once(5.0).with_starts_active(False))
    assert not sp.has_completed()
    assert sp.tick(1.0) == 0 and not sp.has_completed()
    sp.active = True
    assert sp.tick(1.0) == 5 and sp.active and sp.has_completed()
    assert sp.tick(1.0) == 0 and sp.has_completed()
    sp.reset()
    assert sp.active and not sp.has_completed()
    assert sp.tick(1.0) == 5 and sp.has_completed()


def test_rate():
    sp = EffectSpawner(SpawnerSettings.rate(5.0))
    assert sp.tick(1.01) == 5
    assert sp.tick(0.4) == 2
    sp = EffectSpawner(SpawnerSettings.rate(5.0))
    sp.tick(1.01)
    sp.active = False
    assert sp.tick(0.4) == 0
    sp.active = True
    assert sp.tick(0.4) == 2
    sp = EffectSpawner(SpawnerSettings.rate(5.0))
    assert sum(sp.tick(1.0 / 60.0) for _ in range(13)) == 1


def test_burst():
    sp = EffectSpawner(SpawnerSettings.burst(5.0, 2.0))
    assert sp.tick(1.0) == 5
    assert sp.tick(4.0) == 10
    assert sp.tick(0.1) == 0


def test_with_active():
    sp = EffectSpawner(SpawnerSettings.rate(5.0).with_starts_active(False))
    assert not sp.active and sp.tick(1.0) == 0
    sp.active = False
    assert sp.tick(1.0) == 0
    sp.active = True
    assert sp.tick(1.0) == 5


def test_c1_rate_1000_per_second():
    # gpu_tests/single_particle.rs: rate 1000/s at dt=1/60 -> 16,17,17,16,...
    sp = EffectSpawner(SpawnerSettings.rate(1000.0))
    counts = [sp.tick(1.0 / 60.0) for _ in range(60)]
    assert set(counts) <= {16, 17} and 999 <= sum(counts) <= 1000


def _key(asset=1, slab=0, pipe=0, prop=0, parent=0xFFFFFFFF, events=0, cpu=1):
    return BatchKey(asset, slab, pipe, prop, parent, events, cpu)


def test_batcher_merges_same_asset_instances():
    b = Batcher()
    assert b.push(_key(), 0, 0, 10) == 0
    assert b.push(_key(), 1, 512, 5) == -1          # merged
    assert b.push(_key(), 2, 1024, 8) == -1
    assert b.push(_key(asset=2, slab=1), 3, 0, 6) == 1
    infos, prefix, totals = b.finish()
    # headless_batching_tests.rs layout: batch 0 = 3 instances, batch 1 = 1 instance; CPU prefix of spawn counts
    assert [(i.spawner_base, i.base_particle, i.prefix_sum_offset, i.prefix_sum_count) for i in infos] == [(0, 0, 0, 3), (3, 0, 3, 1)]
    assert prefix == [0, 10, 15, 0]
    assert totals == [23, 6]
    assert all(i.total_spawn_count == 0 and i.total_update_count == 0 for i in infos)   # batch.rs:271-278


def test_batcher_never_merges_event_effects_or_different_slabs():
    b = Batcher()
    assert b.push(_key(events=1, cpu=0), 0, 0, 0) == 0
    assert b.push(_key(events=1, cpu=0), 1, 256, 0) == 1     # GPU-event effects stay alone (batch.rs:166-170)
    assert b.push(_key(slab=3), 2, 0, 4) == 2
    assert b.push(_key(slab=4), 3, 0, 4) == 3                # different slab
    assert b.push(_key(slab=4, prop=9), 4, 64, 4) == 4       # different property buffer
    assert b.push(_key(slab=4, prop=9), 5, 128, 4) == -1
    infos, prefix, totals = b.finish()
    assert [i.prefix_sum_count for i in infos] == [1, 1, 1, 1, 2]
    assert prefix == [0, 0, 0, 0, 0, 4] and totals == [0, 0, 4, 4, 8]
    b.clear()
    assert b.finish() == ([], [], [])


def test_effect_sorter_toposort_batches():
    """batch.rs:776-826 `toposort_batches`: children first, then by slab and offset."""
    from bevy_hanabi_b200.spawn import EffectSorter
    s = EffectSorter()
    s.insert(1, 42, 0)                 # some parent effect
    s.insert(2, 5, 30, parent=1)       # a child in a different buffer
    assert s.entities() == [1, 2]
    s.sort()
    assert s.entities() == [2, 1]      # child first, parent after
    s.insert(3, 42, 20, parent=1)      # a child in the same buffer as its parent
    assert s.entities() == [2, 1, 3]   # simply appended
    s.sort()
    assert s.entities() == [2, 3, 1]   # child, other child (same buffer as the parent), finally the parent


def test_effect_sorter_levels_slabs_and_errors():
    from bevy_hanabi_b200.spawn import EffectSorter
    from bevy_hanabi_b200._native import HanabiError
    s = EffectSorter()
    # grandparent 10 <- parent 11 <- child 12; unrelated effects 20, 21 in slab 1, 22 in slab 0
    s.insert(10, 3, 0)
    s.insert(11, 3, 100, parent=10)
    s.insert(12, 4, 0, parent=11)
    s.insert(20, 1, 500)
    s.insert(21, 1, 0)
    s.insert(22, 0, 7)
    s.sort()
    # level 0: 12 (slab 4), 20/21 (slab 1), 22 (slab 0) ordered by (slab, offset); level 1: 11; level 2: 10
    assert s.entities() == [22, 21, 20, 12, 11, 10]
    bad = EffectSorter()
    bad.insert(1, 0, 0, parent=99)
    with pytest.raises(HanabiError):
        bad.sort()
    cyc = EffectSorter()
    cyc.insert(1, 0, 0, parent=2)
    cyc.insert(2, 0, 8, parent=1)
    with pytest.raises(HanabiError):
        cyc.sort()


def test_once_reset():
    """spawn.rs:1163-1173 `test_once_reset`."""
    st = SpawnerSettings.once(5.0)
    assert st.is_once()
    sp = EffectSpawner(st)
    sp.tick(1.0)
    sp.reset()
    assert sp.tick(1.0) == 5


def test_rate_active():
    """spawn.rs:1230-1244 `test_rate_active`: an inactive spawner neither spawns nor accumulates."""
    sp = EffectSpawner(SpawnerSettings.rate(5.0))
    sp.tick(1.01)
    sp.active = False
    assert not sp.active
    assert sp.tick(0.4) == 0
    sp.active = True
    assert sp.active
    assert sp.tick(0.4) == 2


def test_rate_accumulate():
    """spawn.rs:1247-1255 `test_rate_accumulate`: fractions of a particle carry over from tick to tick."""
    sp = EffectSpawner(SpawnerSettings.rate(5.0))
    assert sum(sp.tick(1.0 / 60.0) for _ in range(13)) == 1


def test_uniform_range_is_ordered():
    """spawn.rs:1027-1042 `test_range_*`: CpuValue::range() returns [min, max] whatever the order given; a uniform
    count therefore samples between the two bounds."""
    for bounds in ((1.0, 3.0), (3.0, 1.0)):
        sp = EffectSpawner(SpawnerSettings.once(bounds), rng_seed=5)
        n = sp.tick(1.0)
        assert 1 <= n <= 3


def test_spawner_settings_accessors_of_the_reference():
    """spawn.rs:362-615: with_* / set_* / getters of SpawnerSettings, spawn.rs:723-800 on EffectSpawner."""
    s = SpawnerSettings.rate(5.0)
    assert s.count() == 5.0 and s.spawn_duration() == 1.0 and s.period() == 1.0 and s.cycle_count() == 0 and s.is_forever()
    s = s.with_count((2.0, 4.0)).with_spawn_duration(0.5).with_period((1.0, 3.0)).with_cycle_count(2)
    assert s.count() == (2.0, 4.0) and s.spawn_duration() == 0.5 and s.period() == (1.0, 3.0) and s.cycle_count() == 2
    assert not s.is_once() and not s.is_forever()
    assert s.starts_active() and s.emits_on_start()
    s.set_starts_active(False); s.set_emit_on_start(False)
    assert not s.starts_active() and not s.emits_on_start()
    with pytest.raises(HanabiError, match="infinite bound"):        # spawn.rs:518-545
        s.with_period(float("inf"))
    with pytest.raises(HanabiError, match="infinite bound"):
        s.set_period((1.0, float("inf")))
    assert s.period() == (1.0, 3.0)
    sp = EffectSpawner(SpawnerSettings(3.0, 3.0, 10.0, 2)).with_active(True)
    assert sp.tick(2.0) == 2
    assert (sp.cycle_time(), sp.cycle_spawn_duration(), sp.cycle_period(), sp.cycle_spawn_count(), sp.completed_cycle_count()) == (2.0, 3.0, 10.0, 3.0, 0)
    assert abs(sp.cycle_ratio() - 0.2) < 1e-7
    assert EffectSpawner(SpawnerSettings.once(3.0)).with_active(False).tick(1.0) == 0
