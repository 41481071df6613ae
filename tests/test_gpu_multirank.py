"""Multi-rank parity under pytest: scripts/mgpu_parity.py is spawned with
torch.distributed.run on 2 and 4 GPUs of the box (one rank per GPU, z slabs, NCCL halo exchange, cross-rank
flux reduction) and compared with the single-domain CPU oracle on rank 0; skipped when the box has fewer GPUs.
`bench.py --gpus N` runs the same case and reports it as `extra.parity`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,flags,port", [(2, [], 29611), (2, ["--reference-slabs"], 29612), (2, ["--periodic"], 29613),
                                              (2, ["--kerr"], 29614), (2, ["--nonuniform"], 29615), (4, [], 29616),
                                              (4, ["--periodic"], 29617), (2, ["--bloch"], 29618), (2, ["--blochz"], 29619),
                                              (4, ["--blochz"], 29620), (2, ["--tma"], 29621), (2, ["--tma", "--periodic"], 29622),
                                              (4, ["--tma"], 29623), (4, ["--thin"], 29624), (2, ["--thin", "--tma"], 29625)])
def test_multirank_parity(world, flags, port):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs on the box" % world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "mgpu_parity.py")] + flags,
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "'ok': True" in out.stdout, out.stdout[-2000:]
