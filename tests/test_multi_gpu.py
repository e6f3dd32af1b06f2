"""GPU tests of the sharded mvm's exchange modes, including the fused NVLink-store epilogue."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_exchange_single_rank_equals_mvm():
    """world = 1: the fused path (IPC-exportable result block, epochs, double buffering) returns mvm's bytes."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bench import random_nibbles
    from clover_b200 import containers as cb
    from clover_b200.sharded import ShardedCloverMatrix4
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    rows, cols = 1152, 2048
    rows_pad = rows + (-rows) % 128
    A = ShardedCloverMatrix4(rows, cols, exchange="fused")
    full = cb.CloverMatrix4(rows_pad, cols)
    full.values.copy_(random_nibbles(torch, rows_pad * cols // 2, g, dev))
    full.scales.uniform_(0.05, 4.0, generator=g)
    A.load_shard(full.values, full.scales)
    for _ in range(3):
        v = cb.CloverVector32(cols); v.values.uniform_(-1, 1, generator=g)
        x = cb.CloverVector4(cols); x.quantize(v)
        y, want = cb.CloverVector4(rows_pad), cb.CloverVector4(rows_pad)
        A.mvm(x, y)
        full.mvm(x, want)
        torch.cuda.synchronize()
        assert torch.equal(y.values, want.values)
        assert torch.equal(y.scales.view(torch.int32)[: rows_pad // 64], want.scales.view(torch.int32)[: rows_pad // 64])
    A.close()


def test_exchange_modes_two_gpus():
    """2 ranks over NVLink: fused / allgather / allreduce all reproduce the single-GPU result on every rank."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and "multi-gpu ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
