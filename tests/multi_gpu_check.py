"""Multi-GPU check, launched by torchrun (one rank per GPU): the row-sharded mvm with every exchange mode
("fused" NVLink-store epilogue, "allgather", "allreduce") returns, on EVERY rank, the bytes of the single-GPU
CloverMatrix4::mvm of the whole matrix. Several steps per mode (epochs / double buffering of the fused path),
even and ragged shardings. Prints "multi-gpu ok" on rank 0; any mismatch raises.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import random_nibbles  # noqa: E402
from clover_b200 import containers as cb  # noqa: E402
from clover_b200.sharded import ShardedCloverMatrix4  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)        # a flag protocol that deadlocks must not hold the GPUs
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for rows, cols in [(1024, 2048), (64 * 7 + 64, 1152), (8192, 8192), (4096, 16384 + 128), (19072, 32768)]:   # cols >= 16384: k_m4_mvm_tma2 (the kernel SCALE times)
        rows += (-rows) % 128
        g = torch.Generator(device=dev).manual_seed(7)                    # identical full matrix on every rank
        full = cb.CloverMatrix4(rows, cols)
        full.values.copy_(random_nibbles(torch, rows * cols // 2, g, dev))
        full.scales.uniform_(0.05, 4.0, generator=g)
        xs = []
        for _ in range(5):
            v = cb.CloverVector32(cols); v.values.uniform_(-1, 1, generator=g)
            q = cb.CloverVector4(cols); q.quantize(v); xs.append(q)
        want = []
        for q in xs:
            y = cb.CloverVector4(rows); full.mvm(q, y); want.append((y.values.clone(), y.scales.clone()))
        for mode in ("fused", "fused_sync", "stamped", "allgather", "allreduce"):
            A = ShardedCloverMatrix4(rows, cols, exchange=mode)
            hb = cols // 64
            A.load_shard(full.values[A.row0 * cols // 2:(A.row0 + A.rows_local) * cols // 2],
                         full.scales[(A.row0 // 64) * hb:((A.row0 + A.rows_local) // 64) * hb])
            for step, (q, (wv, ws)) in enumerate(zip(xs, want)):
                if mode in ("fused", "fused_sync", "stamped") and step % 2:          # zero-copy form: a view of the shared result vector
                    y = A.mvm(q) if step % 4 == 1 else A.mvm(q, wait=False)
                    A.wait()
                else:
                    y = cb.CloverVector4(rows)
                    A.mvm(q, y)
                torch.cuda.synchronize()
                assert torch.equal(y.values, wv), (mode, rows, cols, rank, "values")
                assert torch.equal(y.scales.view(torch.int32)[: rows // 64], ws.view(torch.int32)[: rows // 64]), (mode, rows, cols, rank, "scales")
            A.close()
            if rank == 0:
                print(f"multi-gpu check: {rows} x {cols}, exchange={mode}: 5 steps identical to the single-GPU mvm on every rank ({world} ranks)", flush=True)
        dist.barrier()
    # a chain y_{e+1} = A y_e on a square matrix, every step consuming the previous step's result VIEW (assembled from the
    # peers' stamped messages) with no host synchronisation in between; compared with the same chain on one GPU; 3 rounds so
    # that both result buffers and both message areas are re-used, round 2 with explicit wait() calls.
    for n in (2048, 16384 + 128):
        n += (-n) % 128
        g = torch.Generator(device=dev).manual_seed(11)
        full = cb.CloverMatrix4(n, n)
        full.values.copy_(random_nibbles(torch, n * n // 2, g, dev))
        full.scales.uniform_(0.05, 4.0, generator=g)
        v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g)
        x0 = cb.CloverVector4(n); x0.quantize(v)
        steps = 7
        want, cur = [], x0
        for _ in range(steps):
            y = cb.CloverVector4(n); full.mvm(cur, y); want.append(y); cur = y
        A = ShardedCloverMatrix4(n, n, exchange="fused")
        hb = n // 64
        A.load_shard(full.values[A.row0 * n // 2:(A.row0 + A.rows_local) * n // 2],
                     full.scales[(A.row0 // 64) * hb:((A.row0 + A.rows_local) // 64) * hb])
        for rnd in range(3):
            cur = x0
            for e in range(steps):
                if rnd == 2 and e % 3 == 1:
                    cur = A.mvm(cur, wait=False)
                    A.wait()                                # the next step reads this result: unpack first
                else:
                    cur = A.mvm(cur)
            A.wait()
            torch.cuda.synchronize()
            assert torch.equal(cur.values, want[-1].values), ("chain", n, rank, "values")
            assert torch.equal(cur.scales.view(torch.int32)[: n // 64], want[-1].scales.view(torch.int32)[: n // 64]), ("chain", n, rank, "scales")
            dist.barrier()
        A.close()
        if rank == 0:
            print(f"multi-gpu check: {n} x {n}, fused chain of {steps} steps x 3 identical to the single-GPU chain on every rank ({world} ranks)", flush=True)
    if rank == 0:
        print("multi-gpu ok: fused (stamped / flags) / allgather / allreduce == single-GPU mvm on", world, "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
