"""simka_gpu_exchange_and_merge (NCCL inside the library) on 2 GPUs of one box, checked against the oracle by tools/nccl_check.py
under torchrun; skipped on a box with a single GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_nccl_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "nccl_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "native exchange ok" in r.stdout


def test_simka_cli_nb_gpus_2_reproduces_truth(tmp_path, golden_dir):
    """`simka -nb-gpus 2`: one host thread and one context per GPU, NCCL between them; same truth matrices"""
    import gzip
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from simka_b200.build import SIMKA_BIN
    for k, t in ((31, 2), (21, 0)):
        out = tmp_path / f"results_{k}_{t}"
        cmd = [str(SIMKA_BIN), "-in", os.path.join(golden_dir, "example", "simka_input.txt"), "-out", str(out), "-out-tmp", str(tmp_path / f"tmp_{k}_{t}"),
               "-simple-dist", "-complex-dist", "-kmer-size", str(k), "-abundance-min", str(t), "-verbose", "0", "-nb-gpus", "2"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
        checked = 0
        for f in sorted(os.listdir(tdir)):
            assert gzip.open(out / (f + ".gz"), "rt").read() == open(os.path.join(tdir, f)).read(), f
            checked += 1
        assert checked == 20
