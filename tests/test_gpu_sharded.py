"""One logical effect instance sharded over two processes / GPUs, on hardware: see tests/sharded_worker.py.
On a box with one GPU both shards run on cuda:0 (separate processes and contexts); with two or more, on cuda:0 and cuda:1."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_instance_parity(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(ROOT / "tests" / "sharded_worker.py")], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert r.stdout.count(": ok") == world
