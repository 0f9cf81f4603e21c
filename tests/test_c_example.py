"""examples/firework_c_api.c: the boundary is a C ABI, so the whole path must be drivable from plain C (no Python, no
torch): authoring -> lowering -> NVRTC -> spawner tick -> batcher -> hnb_simulate -> draw-args readback."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "build" / "firework_c_api"


def _build():
    (ROOT / "build").mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "firework_c_api.c"), f"-L{ROOT / 'bevy_hanabi_b200'}",
           "-lhanabi_b200", f"-Wl,-rpath,{ROOT / 'bevy_hanabi_b200'}", "-lm", "-o", str(EXE)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_c_example_builds_and_lowers_without_a_gpu():
    """Headers are valid C, the library links from C, and everything up to code generation runs on a CPU-only box;
    creating the context then fails loudly (exit code 3 = HNB_ERR_NO_DEVICE) instead of falling back."""
    import torch
    _build()
    p = subprocess.run([str(EXE), "5"], capture_output=True, text=True, timeout=120)
    assert "lowered 'firework_trails': 32-byte particle records" in p.stdout
    assert "particle.velocity *= max(0.f, (1.f) - ((4.f) * (sim_params.delta_time)));" in p.stdout
    if not torch.cuda.is_available():
        assert p.returncode == 3, p.stderr
        assert "no CUDA device" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu():
    _build()
    p = subprocess.run([str(EXE), "130"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    # bursts of 1000 at t = 0, 1, 2 s; lifetimes 0.8-1.2 s: the first burst is (partly) alive when the second arrives
    assert "ok: 130 frames" in p.stdout
    lines = [l for l in p.stdout.splitlines() if l.startswith("frame")]
    alive = [int(l.rsplit(" ", 1)[1]) for l in lines]
    assert max(alive) >= 1000 and min(alive) >= 0


def test_host_producers_example_runs_without_a_gpu():
    """examples/host_producers_c_api.c: node graph, EffectProperties store and the EffectSimulation clock driven from plain C
    (no context, no device): the new entry points are valid C, link, and produce the tables the runtime consumes."""
    exe = ROOT / "build" / "host_producers_c_api"
    (ROOT / "build").mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "host_producers_c_api.c"), f"-L{ROOT / 'bevy_hanabi_b200'}",
           "-lhanabi_b200", f"-Wl,-rpath,{ROOT / 'bevy_hanabi_b200'}", "-lm", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    out = p.stdout
    assert "graph lowers to: (particle.position) + ((particle.velocity) * (sim_params.delta_time))" in out
    assert "update code: particle.position = (particle.position) + ((particle.velocity) * (sim_params.delta_time));" in out
    assert "properties: 1 stored, changed 1, blob 4 bytes, speed = 7.5" in out
    assert "type mismatch refused: Cannot assign value of type vec3<f32> to property 'speed' of type f32" in out
    assert "frame 2: delta_time 0 " in out and "was_paused 1" in out
    assert "negative speed refused: tried to go back in time" in out
