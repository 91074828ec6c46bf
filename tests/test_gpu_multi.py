"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tools/dist_parity.py under torchrun.  Every rank
compresses its contiguous range of segments, rank 0's gathered stream must equal the oracle's single stream byte for byte
(level 1, level 6, Rle with the halo byte across ranks, HuffmanOnly); the stream is then decoded by inflate_sharded (each
rank one compressed-offset range, Q14 carries across ranks, Adler-32 fold against the trailer)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_deflate_and_inflate_parity(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join("tools", "dist_parity.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIFFERENT" not in r.stdout and r.stdout.count("IDENTICAL") == 8
