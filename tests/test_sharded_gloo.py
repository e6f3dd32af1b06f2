"""Host-side logic of the row-sharded mvm on CPU: world_size 2 (and 3, uneven) over gloo.

Each rank computes the fp32 results of ITS rows with the oracle, the collective of clover_b200.sharded
makes the vector whole, every rank re-quantizes - and must end up with exactly the single-process
CloverVector4 (the sharding must be invisible in the result bits)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clover_b200.sharded import exchange_fp32, shard_rows


def test_shard_rows_partition():
    for rows in (128, 256, 1280, 65536):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_rows(rows, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == rows
            for (a, n), (b, _) in zip(parts, parts[1:]):
                assert a + n == b
            assert all(a % 64 == 0 and n % 64 == 0 for a, n in parts)
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 64


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle, pad_matrix
        orc = Oracle()
        rows, cols = 384, 256
        st = orc.xs_init()
        a = pad_matrix(orc.fill_floats(rows * cols, -1.0, 1.0, st)[: rows * cols].reshape(rows, cols))
        xvec = orc.fill_floats(cols, -1.0, 1.0, st)
        mv, ms = orc.m4_quantize(a)
        xv, xs = orc.v4_quantize(xvec, cols)
        want_v, want_s, want32 = orc.m4_mvm(mv, ms, rows, cols, xv, xs, want_f32=True)

        row0, nloc = shard_rows(rows, world, rank)
        sizes = [shard_rows(rows, world, r)[1] for r in range(world)]
        y32 = torch.zeros(rows, dtype=torch.float32)
        if nloc:
            lv = mv[row0 * cols // 2:(row0 + nloc) * cols // 2]
            ls = ms[(row0 // 64) * (cols // 64):((row0 + nloc) // 64) * (cols // 64)]
            # mvm on the local rows only; 64-row shards are padded to the oracle's x128 only in the test harness
            pad = (-nloc) % 128
            lv = np.concatenate([lv, np.zeros(pad * cols // 2, np.int8)])
            ls = np.concatenate([ls, np.ones((pad // 64) * (cols // 64), np.float32)])
            _, _, part = orc.m4_mvm(lv, ls, nloc + pad, cols, xv, xs, want_f32=True)
            y32[row0:row0 + nloc] = torch.from_numpy(part[:nloc].copy())
        exchange_fp32(y32, row0, nloc, sizes, mode)
        got32 = y32.numpy()
        assert np.array_equal(got32.view(np.uint32), want32.view(np.uint32)), "fp32 output after the collective"
        # every rank re-quantizes the full vector: same bytes as the unsharded mvm
        # (with rounding disabled the mvm tail is arithmetically the block-64 vector quantizer)
        gv, gs = orc.v4_quantize(np.ascontiguousarray(got32), rows)
        assert np.array_equal(gv, want_v) and np.array_equal(gs.view(np.uint32), want_s.view(np.uint32))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "allreduce"), (2, "allgather"), (3, "allgather"), (3, "allreduce")])
def test_sharded_mvm_over_gloo(tmp_path, world, mode):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_peer_block_layout_is_aligned_and_disjoint():
    """The IPC-shared block of the fused exchange: two result vectors, flags, ticket, two message areas of the stamped
    exchange (9 x 8 bytes per 64-row block of the WHOLE vector), started words - 256-byte aligned, non-overlapping."""
    from clover_b200.sharded import ShardedCloverMatrix4, shard_rows
    for rows in (128, 1152, 65536):
        for world in (1, 2, 3, 8):
            lay = ShardedCloverMatrix4.peer_block_layout(rows, world)
            spans = [(lay["yv"][0], rows // 2), (lay["yv"][1], rows // 2), (lay["ys"][0], rows // 64 * 4), (lay["ys"][1], rows // 64 * 4),
                     (lay["flags"], 4 * world), (lay["ticket"], 12), (lay["msg"][0], rows // 64 * 72), (lay["msg"][1], rows // 64 * 72),
                     (lay["started"], 4 * world)]
            assert all(off % 256 == 0 for off, _ in spans)
            spans.sort()
            assert all(a + n <= b for (a, n), (b, _) in zip(spans, spans[1:]))
            assert spans[-1][0] + spans[-1][1] <= lay["bytes"]
            # shards: whole 64-row blocks, contiguous, covering every block exactly once
            pos = 0
            for r in range(world):
                row0, n = shard_rows(rows, world, r)
                assert row0 == pos and n % 64 == 0
                pos += n
            assert pos == rows
