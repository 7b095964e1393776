"""Multi-rank parity that runs where the driver runs: ONE GPU.  Two (and three) processes share cuda:0 and exchange
halos and reductions through the library's own peer-to-peer transport (B200_TRANSPORT=p2p, cudaIpc-mapped buffers) -
the transport `bench.py --gpus N` uses over NVLink - so SURVEY rows a16 (processor-patch halo), a17 (global sums) and
the all-reduce leg of a20 have a green record that does not need a second device."""
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_selfpeer_multirank_parity(world):
    uid = os.urandom(128).hex()
    with tempfile.TemporaryDirectory() as work:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "scripts", "selfpeer_parity.py"), work, str(rk), str(world), uid, "1", "3"],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for rk in range(world)]
        outs = []
        try:
            for p in procs:
                outs.append(p.communicate(timeout=900)[0])
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
        assert all(p.returncode == 0 for p in procs), "\n".join(f"[rank {i} rc {p.returncode}] " + o[-2500:] for i, (p, o) in enumerate(zip(procs, outs)))
        assert "amul_bit_exact=True" in outs[0] and "zone_allreduce_bit_exact=True" in outs[0], outs[0][-2000:]
