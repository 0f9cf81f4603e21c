import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import bevy_hanabi_b200 as hb
from bevy_hanabi_b200 import recipes
from oracle import c_oracle as O
from tests.helpers import GpuWorld, Instance, RefWorld
from tests.test_gpu_fullsize import _oracle_run, _gpu_run
orc = O.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ctx = hb.Context(0)
for steps in (1, 2, 3, 12):
    w = _oracle_run(orc, n, 1234, 0.02, 0.3, steps, 1/60); g = _gpu_run(ctx, n, 1234, 0.02, 0.3, steps, 1/60)
    print(steps, {k: (w[k] == g[k]) for k in w}, w["alive"], g["alive"])
# detailed diff with full buffers
ref = RefWorld(n, 8, [Instance(0, n, alive=n, seed=42)])
orc.orc_fill_c5(O.ptr(ref.particles.view(np.float32)), O.ptr(ref.indirect), 0, n, 1234, 0.02, 0.3)
gpu = GpuWorld(ctx, ref, recipes.c5_lowered())
k = (C.c_float * 4)(0.0, -9.8, 0.0, 0.5)
for step in range(4):
    ref.oracle_frame(orc, orc.orc_body_update_c5(), k)
    gpu.frame()
    got = gpu.pull()
    for name, a, b in (("particles", got["particles"], ref.particles), ("indirect", got["indirect"], ref.indirect), ("metadata", got["metadata"], ref.metadata_rows()), ("draw", got["draw"], ref.draw)):
        bad = np.argwhere(a != b)
        print(step, name, "mismatches", len(bad), bad[:5].tolist(), (a[tuple(bad[0])], b[tuple(bad[0])]) if len(bad) else "")
