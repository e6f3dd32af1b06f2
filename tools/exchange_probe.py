"""Multi-GPU exchange probe (torchrun, one rank per GPU): per-rank step time of the sharded C3 mvm with the flag-synchronised
fused exchange, the stamped one (with one unpack at the end and with an unpack per step) and the bare shard kernel, all in
one process.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/exchange_probe.py [steps=200]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from clover_b200 import containers as cb
from clover_b200.sharded import ShardedCloverMatrix4
from bench import random_nibbles


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rows = cols = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = cb.CloverVector4(cols)
    x.values.copy_(random_nibbles(torch, cols // 2, torch.Generator(device=dev).manual_seed(5), dev))
    x.scales.uniform_(0.25, 1.0, generator=torch.Generator(device=dev).manual_seed(6))
    res = {}
    mats = {}
    for mode in ("fused_sync", "stamped"):
        A = ShardedCloverMatrix4(rows, cols, exchange=mode)
        if mode == "fused_sync":
            A.local.values[: A.rows_local * cols // 2].copy_(random_nibbles(torch, A.rows_local * cols // 2, g, dev))
            A.local.scales.uniform_(0.25, 1.0, generator=g)
        else:
            A.local = mats["fused_sync"].local          # same shard
        mats[mode] = A

    def timed(name, fn, fin=lambda: None):
        for _ in range(5):
            fn()
        fin()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        fin()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        res[name] = e0.elapsed_time(e1) / steps * 1e3

    S, T = mats["fused_sync"], mats["stamped"]
    y = cb.CloverVector4(rows)
    for rep in range(1):
        timed(f"flags_{rep}", lambda: S.mvm(x))
        timed(f"stamped_{rep}", lambda: T.mvm(x, wait=False), T.wait)
        timed(f"stamped_unpack_each_{rep}", lambda: T.mvm(x, wait=True))
        ys_, yt_ = S.mvm(x, wait=True), T.mvm(x, wait=True)
        torch.cuda.synchronize()
        res[f"stamped_equals_sync_{rep}"] = float(torch.equal(ys_.values, yt_.values) and torch.equal(ys_.scales.view(torch.int32)[: rows // 64], yt_.scales.view(torch.int32)[: rows // 64]))
    import ctypes as C
    import clover_b200
    def shard():
        clover_b200.call("clover_m4_mvm_shard", C.c_void_p(S.local.values.data_ptr()), C.c_void_p(S.local.scales.data_ptr()),
                         C.c_uint64(S.rows_local), C.c_uint64(cols), C.c_uint64(S.row0), C.c_void_p(x.values.data_ptr()),
                         C.c_void_p(x.scales.data_ptr()), None, C.c_void_p(y.values.data_ptr()), C.c_void_p(y.scales.data_ptr()), None,
                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    timed("shard_kernel_only", shard)
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        for k in res:
            print(k, "us per step by rank:", [round(r[k], 1) for r in allres], flush=True)
    S.close(); T.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
