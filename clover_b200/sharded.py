"""Row-sharded multi-GPU mvm - one process per GPU, torch.distributed/NCCL for the plumbing.

The reference has no distributed code; this is the multi-GPU form of ``CloverMatrix4::mvm`` that
BASELINE.json's north_star asks for: rows are sharded in whole 64-row blocks (so every re-quantization
block is local to one rank), the product vector is replicated, every rank computes the fp32 results of its
own rows, and ONE collective on the fp32 output makes the full vector visible everywhere:

  * ``exchange="allreduce"`` (north_star): ncclAllReduce(sum) over a zero-initialised full-length fp32
    vector in which each rank filled only its slice;
  * ``exchange="allgather"``: ncclAllGather of the rows/G slices - same result, 1/G of the traffic.

After the exchange every rank re-quantizes the full fp32 vector with the mvm epilogue
(include/CloverMatrix4.h:925-1080), so all ranks hold the same CloverVector4 the single-GPU call returns -
bit-for-bit, because the fp32 row results do not depend on the sharding.

  * ``exchange="fused"``: no collective call at all. Re-quantization blocks are 64 rows and shards are whole
    blocks, so each rank's GEMV kernel re-quantizes its own blocks and its epilogue stores them (36 bytes per
    block) directly into every peer's memory over NVLink (peer memory mapped with CUDA IPC). One kernel per step
    and rank; the same bytes as the single-GPU call. Two protocols, chosen per call by what was measured
    (tools/exchange_probe.py, C3 sharded, us per step on one HGX box):

      - *stamped* (``clover_m4_mvm_shard_stamped``): every 32-bit word of a finished block travels to every peer
        in ONE 8-byte store {word, epoch} into a message area - no ordering between stores, so no system-scope
        fence per CTA, no flags; the kernel just ends. ``wait()`` (``clover_m4_shard_stamped_unpack``) polls the
        stamps and writes the words into the reference layout in front of whatever reads the result.
      - *flags* (``clover_m4_mvm_shard_fused``): plain stores into the peers' result vectors, a system-scope fence
        per CTA, one flag word per peer, and the kernel waits for the peers' flags at its end.

                                    2 GPUs    4 GPUs    8 GPUs
        flags                        180.8     101.5      63.6
        stamped, wait=False          178.7      95.0      50.8     (one unpack at the end)
        stamped + unpack per step    184.5     101.5      55.8
        shard kernel alone           170       89.5       45.7

    ``mvm(x, wait=False)`` therefore always uses the stamped kernel (call ``wait()`` before the result is read);
    ``mvm(x)`` (complete when it returns in stream order) uses stamped + unpack from 4 ranks on, flags below.
    ``exchange="stamped"`` / ``"fused_sync"`` force either.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import call
from .containers import CloverMatrix4, CloverSizeError, CloverVector4, _ptr, _stream


def shard_rows(rows: int, world: int, rank: int):
    """Contiguous ranges of 64-row blocks, remainder spread over the first ranks."""
    blocks = rows // 64
    base, extra = divmod(blocks, world)
    b0 = rank * base + min(rank, extra)
    nb = base + (1 if rank < extra else 0)
    return b0 * 64, nb * 64


def exchange_fp32(y32: torch.Tensor, row0: int, rows_local: int, sizes, mode: str = "allreduce", group=None) -> None:
    """The ONE collective of the sharded mvm, on the full-length fp32 output (any device/backend).

    On entry ``y32[row0:row0+rows_local]`` holds this rank's results; for ``allreduce`` the rest must be
    zero. On exit every rank holds all rows. ``sizes`` = rows per rank, in rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return
    if mode == "allreduce":
        dist.all_reduce(y32, op=dist.ReduceOp.SUM, group=group)
    elif mode == "allgather":
        mine = y32[row0:row0 + rows_local].clone()
        if len(set(sizes)) == 1 and y32.is_cuda:
            dist.all_gather_into_tensor(y32, mine, group=group)
        else:
            dist.all_gather(list(torch.split(y32, list(sizes))), mine, group=group)
    else:
        raise ValueError(f"unknown exchange {mode!r}")


class ShardedCloverMatrix4:
    """This rank's rows [row0, row0 + rows_local) of a (rows x cols) CloverMatrix4."""

    def __init__(self, rows: int, cols: int, group=None, device=None, exchange: str = "allreduce"):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows = rows + (-rows) % 128
        self.cols = cols + (-cols) % 128
        self.row0, self.rows_local = shard_rows(self.rows, self.world, self.rank)
        self.exchange = exchange
        self.device = device
        if self.rows_local:
            # a CloverMatrix4 of rows_local x cols (rows_local is a multiple of 64; pad to 128 only in storage)
            self.local = CloverMatrix4(self.rows_local + (-self.rows_local) % 128, self.cols, device=device)
        else:
            self.local = None
        dev = self.local.values.device if self.local is not None else torch.device("cuda", torch.cuda.current_device())
        self.y32 = torch.zeros(self.rows, dtype=torch.float32, device=dev)
        sizes = [shard_rows(self.rows, self.world, r)[1] for r in range(self.world)]
        self._even = len(set(sizes)) == 1
        self._sizes = sizes
        self.key = None
        self._peer = None
        self._pending_wait = False                     # the last step was stamped and its messages are not unpacked yet
        if exchange in ("fused", "fused_sync", "stamped"):           # fused = by measurement; the other two force a protocol
            self._setup_peer_memory()

    # ---- fused exchange: one IPC-shared block per rank = [values x2 | scales x2 | flags | ticket] ------------------
    @staticmethod
    def peer_block_layout(rows: int, world: int):
        """Byte offsets inside a rank's shared block: two result buffers (alternating by epoch), flags, ticket."""
        al = lambda n: (n + 255) // 256 * 256
        vb, sb = al(rows // 2), al(rows // 64 * 4)
        off = {"yv": (0, vb), "ys": (2 * vb, 2 * vb + sb), "flags": 2 * vb + 2 * sb}
        off["ticket"] = off["flags"] + al(4 * world)
        # stamped exchange: two message areas (9 x 8 bytes per 64-row block of the whole vector) and the started words
        mb = al(rows // 64 * 72)
        off["msg"] = (off["ticket"] + 256, off["ticket"] + 256 + mb)
        off["started"] = off["msg"][1] + mb
        off["bytes"] = off["started"] + al(4 * world)
        return off

    def _setup_peer_memory(self) -> None:
        from ._lib import lib
        if self.world > 8:
            raise ValueError("fused exchange: one node, at most 8 ranks")
        if self.rows_local == 0:
            raise ValueError("fused exchange: every rank needs at least one 64-row block")
        lay = self.peer_block_layout(self.rows, self.world)
        base = C.c_void_p()
        call("clover_malloc", C.byref(base), C.c_size_t(lay["bytes"]))
        call("clover_memset", base, 0, C.c_size_t(lay["bytes"]), None)
        call("clover_stream_sync", None)
        handle = (C.c_ubyte * 64)()
        call("clover_ipc_export", base, handle)
        bases = [None] * self.world
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=self.group)
            for p, h in enumerate(handles):
                if p == self.rank:
                    bases[p] = base.value
                else:
                    ptr = C.c_void_p()
                    call("clover_ipc_import", (C.c_ubyte * 64).from_buffer_copy(h), C.byref(ptr))
                    bases[p] = ptr.value
            dist.barrier(group=self.group)
        else:
            bases[0] = base.value
        arr = lambda f: (C.c_void_p * self.world)(*[f(b) for b in bases])

        class _Raw:                      # device memory of this rank's block as a torch tensor (zero copy)
            def __init__(self, ptr, nbytes, typestr):
                item = 1 if typestr == "|i1" else 4
                self.__cuda_array_interface__ = {"shape": (nbytes // item,), "typestr": typestr, "data": (ptr, False), "version": 2}

        dev = self.local.values.device
        views = []
        for k in (0, 1):                 # the two result vectors as borrowing CloverVector4 views (CloverVector4.h:114-119)
            v = torch.as_tensor(_Raw(base.value + lay["yv"][k], self.rows // 2, "|i1"), device=dev)
            sc = torch.as_tensor(_Raw(base.value + lay["ys"][k], self.rows // 64 * 4, "<f4"), device=dev)
            views.append(CloverVector4(self.rows, v, sc, device=dev))
        self._peer = {
            "views": views,
            "lay": lay, "base": base.value, "bases": bases, "epoch": 0,
            "yv": [arr(lambda b, k=k: b + lay["yv"][k]) for k in (0, 1)],
            "ys": [arr(lambda b, k=k: b + lay["ys"][k]) for k in (0, 1)],
            "flags": arr(lambda b: b + lay["flags"]),
            "ticket": C.c_void_p(base.value + lay["ticket"]),
            "msg": [arr(lambda b, k=k: b + lay["msg"][k]) for k in (0, 1)],
            "started": arr(lambda b: b + lay["started"]),
        }

    def close(self) -> None:
        """Unmap the peers' blocks and free this rank's (collective: every rank calls it)."""
        if self._peer is None:
            return
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        for p, b in enumerate(self._peer["bases"]):
            if p != self.rank:
                call("clover_ipc_close", C.c_void_p(b))
        if self.world > 1:
            dist.barrier(group=self.group)
        call("clover_free", C.c_void_p(self._peer["base"]))
        self._peer = None

    def wait(self) -> None:
        """Fused exchange, after ``mvm(x, wait=False)``: enqueue the unpack of the LAST step's messages on the current
        stream. When it has completed in stream order, the whole result of that step is present on this rank in the
        reference layout (no-op otherwise; the NCCL modes' collectives are stream-ordered already)."""
        pr = self._peer
        if pr is None or self.world == 1 or pr["epoch"] == 0 or not self._pending_wait:
            return
        self._pending_wait = False
        k = pr["epoch"] & 1
        call("clover_m4_shard_stamped_unpack", C.c_void_p(pr["base"] + pr["lay"]["msg"][k]), C.c_uint64(self.rows),
             C.c_uint64(self.row0), C.c_uint64(self.rows_local), C.c_uint32(pr["epoch"]),
             C.c_void_p(pr["base"] + pr["lay"]["yv"][k]), C.c_void_p(pr["base"] + pr["lay"]["ys"][k]), _stream())

    def _mvm_fused(self, x: CloverVector4, y, key_ptr, wait: bool):
        """y = None: returns a CloverVector4 VIEW of the shared result buffer of this step - no copy at all. The view is
        valid until this rank issues its NEXT mvm (a faster peer may then already be storing the step after that into the
        same buffer); otherwise the result is copied into the caller's vector. wait = False: the result is complete only
        after ``wait()``."""
        pr = self._peer
        stamped = self.exchange == "stamped" or (self.exchange == "fused" and (not wait or self.world >= 4))
        if stamped:
            return self._mvm_stamped(x, y, key_ptr, wait)
        self._pending_wait = False           # a stamped step nobody waited for is dropped: its buffers are simply re-used
        pr["epoch"] += 1
        k = pr["epoch"] & 1
        call("clover_m4_mvm_shard_fused", _ptr(self.local.values), _ptr(self.local.scales), C.c_uint64(self.rows_local),
             C.c_uint64(self.cols), C.c_uint64(self.row0), _ptr(x.values), _ptr(x.scales), pr["yv"][k], pr["ys"][k],
             pr["flags"], pr["ticket"], self.world, self.rank, C.c_uint32(pr["epoch"]), key_ptr, _stream())
        lay = pr["lay"]
        if key_ptr is not None:      # the kernel read the key at each block's global position; advance it like the reference
            call("clover_prng_skip", key_ptr, C.c_uint64(2 * (self.rows // 64)))
        if y is None:
            return pr["views"][k]
        call("clover_copy_d2d", _ptr(y.values), C.c_void_p(pr["base"] + lay["yv"][k]), C.c_size_t(self.rows // 2), _stream())
        call("clover_copy_d2d", _ptr(y.scales), C.c_void_p(pr["base"] + lay["ys"][k]), C.c_size_t(self.rows // 64 * 4), _stream())
        return y

    def _mvm_stamped(self, x: CloverVector4, y, key_ptr, wait: bool):
        """The stamped exchange: kernel (messages to every peer, own blocks in place), then - unless wait is False - the
        unpack pass. ``wait()`` always unpacks the LAST call; a call whose messages were never unpacked is simply
        dropped (its message area is re-used two calls later, after every peer has started that call)."""
        pr = self._peer
        lay = pr["lay"]
        pr["epoch"] += 1
        k = pr["epoch"] & 1
        call("clover_m4_mvm_shard_stamped", _ptr(self.local.values), _ptr(self.local.scales), C.c_uint64(self.rows_local),
             C.c_uint64(self.cols), C.c_uint64(self.row0), _ptr(x.values), _ptr(x.scales),
             C.c_void_p(pr["base"] + lay["yv"][k]), C.c_void_p(pr["base"] + lay["ys"][k]), pr["msg"][k], pr["started"],
             self.world, self.rank, C.c_uint32(pr["epoch"]), key_ptr, _stream())
        if key_ptr is not None:
            call("clover_prng_skip", key_ptr, C.c_uint64(2 * (self.rows // 64)))
        self._pending_wait = True
        if wait or y is not None:
            self.wait()
        if y is None:
            return pr["views"][k]
        call("clover_copy_d2d", _ptr(y.values), C.c_void_p(pr["base"] + lay["yv"][k]), C.c_size_t(self.rows // 2), _stream())
        call("clover_copy_d2d", _ptr(y.scales), C.c_void_p(pr["base"] + lay["ys"][k]), C.c_size_t(self.rows // 64 * 4), _stream())
        return y

    def load_shard(self, values, scales) -> None:
        """values/scales of this rank's rows in the reference layout (rows_local*cols/2 bytes, tile-row scales)."""
        nb = self.rows_local * self.cols // 2
        ns = (self.rows_local // 64) * (self.cols // 64)
        self.local.values[:nb].copy_(torch.as_tensor(values).view(torch.int8).reshape(-1)[:nb])
        self.local.scales[:ns].copy_(torch.as_tensor(scales).reshape(-1)[:ns])

    def mvm(self, x: CloverVector4, y: CloverVector4 = None, wait: bool = True):
        if x.size() != self.cols or (y is not None and y.size_pad() != self.rows):
            raise CloverSizeError("MVM can not be performed.")
        fused = self.exchange.startswith("fused") or self.exchange == "stamped"
        if y is None and not fused:
            y = CloverVector4(self.rows, device=self.device)
        key_ptr = None if self.key is None else self.key.ctypes.data_as(C.c_void_p)
        if fused:
            return self._mvm_fused(x, y, key_ptr, wait)
        if self.exchange == "allreduce":
            self.y32.zero_()
        if self.rows_local:
            call("clover_m4_mvm_shard", _ptr(self.local.values), _ptr(self.local.scales), C.c_uint64(self.rows_local),
                 C.c_uint64(self.cols), C.c_uint64(self.row0), _ptr(x.values), _ptr(x.scales), _ptr(self.y32),
                 None, None, None, _stream())
        exchange_fp32(self.y32, self.row0, self.rows_local, self._sizes, self.exchange, self.group)
        call("clover_v4_requantize_mvm", _ptr(self.y32), C.c_uint64(self.rows), _ptr(y.values), _ptr(y.scales),
             key_ptr, _stream())
        return y
