"""N-rank vs 1-rank prefill on real hardware (torchrun + NCCL + peer-memory exchange), the check SURVEY.md 4(c) asks for in place of
the reference's Gather.forward (lmm/dattn/sequence_parallel/all_to_all.py:361, split.py:72-93).

Spawns ``torchrun --nproc-per-node N tests/dist_worker.py`` for every N in {2, 4, 8} that the box has GPUs for and asserts, for both
exchange modes (peer-memory stores + flag wait; NCCL all-gather):  logits bit-equal on all ranks, rel-L2 <= 1e-2 against the SAME
engine run as one rank on the same box, identical argmax wherever the 1-rank top-2 margin exceeds 4x the max-abs deviation.
On a 1-GPU box these tests skip; tests/test_engine_gpu.py::test_multirank_text_path_lockstep_equals_single_rank runs the same engine
code for fake ranks on one device."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world: int, case: str, tmp_path):
    out = tmp_path / f"dist_{case}_{world}.json"
    port = 29600 + (os.getpid() + world * 7) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), "--case", case, "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(out.read_text())


@pytest.mark.parametrize("case", ["mini", "c3cut"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nrank_equals_1rank(world, case, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    res = _run(world, case, tmp_path)
    for mode in ("p2p", "nccl"):
        m = res[mode]
        assert m["ranks_bit_equal"], (mode, m)
        assert m["rel_l2"] <= 1e-2, (mode, m)
        assert m["argmax_equal_on_decisive"] and m["decisive_positions"] > 0, (mode, m)
    assert res["nccl"]["used"] == "nccl"
    if case == "c3cut":
        assert res["image_hw"] == [10, 10]
