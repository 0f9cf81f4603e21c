"""The benchmark workload at full size through the EMULATED update kernel (tests/kernel_emu.py): BASELINE.json's C5,
64 Mi particles in one instance, two update steps (the second with ~10 % deaths), compared word for word with the C
oracle — every alive / dead list entry, counter, draw-indirect count and particle word. The GPU suite compares the same
size through whole-state checksums (tests/test_gpu_fullsize.py); this is the exhaustive CPU counterpart (~3 min, ~12 GB).

    python tools/emu_c5_fullsize.py [particles] [instances]     # instances > 1: BASELINE's C4 topology (equal slices of one
                                                                # slab in ONE batch, each filled to a random level)
                                                                # tiles (3552) at the tail of every instance in 128-row tiles
"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from bevy_hanabi_b200 import recipes
from oracle import c_oracle
from tests.helpers import Instance, RefWorld
from tests.kernel_emu import EmuWorld
orc=c_oracle.load()
n=int(sys.argv[1]) if len(sys.argv)>1 else (1<<26)
t=time.time()
E=int(sys.argv[2]) if len(sys.argv)>2 else 1
rng=np.random.default_rng(7)
cap=n//E
insts=[Instance(i*cap,cap,alive=(cap if E==1 else int(rng.integers(0,cap+1))),seed=42+i) for i in range(E)]
ref=RefWorld(n,8,insts)
p=ref.particles.view(np.float32)
blk=1<<22
for s in range(0,n,blk):   # rows beyond an instance's alive count are never read
    m=min(blk,n-s)
    p[s:s+m,0:3]=rng.uniform(-1,1,(m,3)); p[s:s+m,4:7]=rng.uniform(-1,1,(m,3)); p[s:s+m,7]=rng.uniform(0.02,0.15,m); p[s:s+m,3]=0
print("world built", f"{time.time()-t:.0f} s", flush=True)
emu=EmuWorld(ref, recipes.c5_lowered(), chunks=4, update_ctas=3)
print("emu built", f"{time.time()-t:.0f} s", flush=True)
k=(C.c_float*4)(0.0,-9.8,0.0,0.5)
for step in range(2):
    ref.oracle_frame(orc, orc.orc_body_update_c5(), k)
    print("oracle frame", step, f"{time.time()-t:.0f} s", flush=True)
    emu.frame_step(orc, ref.sim, [0]*E, [42+i for i in range(E)])
    print("emulated frame", step, f"{time.time()-t:.0f} s", flush=True)
    got=emu.pull()
    assert np.array_equal(got["metadata"], ref.metadata_rows())
    assert np.array_equal(got["draw"], ref.draw)
    assert np.array_equal(got["indirect"], ref.indirect), "lists"
    assert np.array_equal(got["particles"], ref.particles), "particles"
    print("step", step, "alive", sum(int(m.alive_count) for m in ref.metadata), "exact", f"{time.time()-t:.0f} s", flush=True)
    del got
print(f"C5 at {n} slots in {E} instance(s): {2} update steps through the emulated hnb_update (3 CTAs, 512-row tiles), every list entry, counter and particle word equal to the oracle")
