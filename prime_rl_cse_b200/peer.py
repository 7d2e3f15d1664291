"""Peer-mapped device buffers for the fused compute+collective epilogues (include/prl_b200.h: prl_ipc_*).

One fp32 [rows, cols] buffer per rank, each a cudaMalloc allocation of its own that is exported over CUDA IPC and mapped by
every other rank of the node under ITS device (peer access over NVLink).  The GEMM engine's TMA reduce-add epilogue then adds
output tiles straight into the owner's buffer: the reduce-scatter of d_hidden of the vocab-sharded LM head
(`parallel.VocabShardedLMHead`), and the accumulation of d_W_lm directly in the reduce-scattered (FSDP) layout
(`ShardedGradAccumulator`).  Every step that can fail on one rank only is followed by an agreement, so that ALL ranks raise
together instead of leaving the others inside a collective."""

from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


def _agree(ok: bool, dev, group) -> bool:
    """True iff `ok` on every rank of the group"""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class PeerBuffer:
    """rows x cols fp32 on every rank (rows may differ per rank); `local` is this rank's buffer as a torch tensor, `ptrs` a
    ctypes array of all ranks' device pointers (own one included) to hand to the C ABI."""

    def __init__(self, rows: int, cols: int, device, group: dist.ProcessGroup | None = None):
        lib = _lib.load()
        self.group, self.rank, self.world = group, dist.get_rank(group), dist.get_world_size(group)
        self.rows, self.cols = int(rows), int(cols)
        nbytes = max(self.rows, 1) * self.cols * 4
        base, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        err = None
        try:
            _lib.check(lib.prl_ipc_alloc(nbytes, ctypes.byref(base), handle), "prl_ipc_alloc")
        except Exception as e:  # noqa: BLE001
            err = e
        if not _agree(err is None, device, group):
            if base.value:
                lib.prl_ipc_free(base)
            raise RuntimeError(f"peer buffers unavailable: prl_ipc_alloc failed on a rank ({err})")
        shape = (max(self.rows, 1), self.cols)

        class _Buf:  # zero-copy view for torch (CUDA array interface v2)
            __cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(base.value), False), "version": 2}

        self.local = torch.as_tensor(_Buf(), device=device)[: self.rows]
        self.local.zero_()
        infos = [None] * self.world
        dist.all_gather_object(infos, bytes(handle.raw), group=group)
        ptrs, self._opened = [], []
        try:
            for j, h in enumerate(infos):
                if j == self.rank:
                    ptrs.append(int(base.value))
                    continue
                peer = ctypes.c_void_p()
                _lib.check(lib.prl_ipc_open(ctypes.c_char_p(h), ctypes.byref(peer)), "prl_ipc_open")
                self._opened.append(peer.value)
                ptrs.append(peer.value)
        except Exception as e:  # noqa: BLE001
            err = e
        if not _agree(err is None, device, group):
            for p_ in self._opened:
                lib.prl_ipc_close(p_)
            dist.barrier(group=group)
            self.local = None
            lib.prl_ipc_free(base)
            raise RuntimeError(f"peer buffers unavailable: mapping a peer's buffer failed on a rank ({err})")
        self._own = int(base.value)
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        dist.barrier(group=group)  # every rank has mapped every buffer before anyone uses them

    def barrier(self) -> None:
        """stream-ordered cross-rank barrier (a 4-byte all-reduce): everything the ranks queued before it has completed --
        including their epilogues' adds into this rank's buffer -- before anything queued after it starts"""
        dist.all_reduce(self._flag, group=self.group)

    def close(self) -> None:
        if self._own is None:
            return
        lib = _lib.load()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p_ in self._opened:
            lib.prl_ipc_close(p_)
        dist.barrier(group=self.group)  # nobody maps the buffer any more
        own, self._own, self.local = self._own, None, None
        lib.prl_ipc_free(own)


class ShardedGradAccumulator:
    """d_W_lm delivered directly in the reduce-scattered (FSDP2 dim-0) layout across the ranks of a node: rank j owns rows
    [j * per, (j + 1) * per) of the [V, H] gradient, per = ceil(V / world), as an fp32 PeerBuffer that the d_weight GEMM epilogues
    of ALL ranks add into (prl_lmhead_bwd_dw_scatter).  Pass it as `main_grad` to `fused_lm_head_logprobs_entropy` /
    `rl_micro_step` / `FusedLMHead` (recompute backward).  Per optimizer step:

        acc.begin_step(n)           # n = micro-batches of this step; zero the own shard; barrier
        ... n micro-batches ...     # backward, no collective
        shard = acc.finish_step()   # barrier; the own shard now holds sum over ranks of their gradients for these rows

    Two schedules.  Default: the first n - 1 micro-batches accumulate into a LOCAL fp32 [V, H] buffer (TMA reduce-add, as without
    the accumulator) and the LAST one sends (local total + its own tiles) -- the NVLink traffic of one reduce-scatter per step,
    hidden behind that micro-batch's GEMMs.  `every_micro_batch=True`: every micro-batch adds straight into the owners' shards and
    no rank holds a full [V, H] gradient (2.49 GB -> 0.31 GB at 8 ranks for Qwen3) at n times the NVLink traffic (measured: 1.2 %
    slower at 8 ranks).  The order in which the ranks' addends reach a shard is not fixed: reproducible to fp32 rounding, not
    bit for bit."""

    def __init__(self, vocab_size: int, hidden_size: int, device, group: dist.ProcessGroup | None = None, *, every_micro_batch: bool = False,
                 local: torch.Tensor | None = None):
        self.V, self.H = int(vocab_size), int(hidden_size)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8:
            raise ValueError("ShardedGradAccumulator: at most 8 ranks (one NVLink domain)")
        self.rows_per_peer = -(-self.V // world)
        if self.rows_per_peer < 32:
            raise ValueError("ShardedGradAccumulator: fewer than 32 vocab rows per rank")
        v0 = min(rank * self.rows_per_peer, self.V)
        self.v0, self.v1 = v0, min(v0 + self.rows_per_peer, self.V)
        self.buf = PeerBuffer(self.v1 - self.v0, self.H, device, group)
        self.rank, self.world = rank, world
        self.every_micro_batch = bool(every_micro_batch)
        self.local = None
        if not self.every_micro_batch:
            if local is not None and (local.dtype != torch.float32 or tuple(local.shape) != (self.V, self.H) or not local.is_contiguous()):
                raise ValueError("local accumulator must be a contiguous fp32 [V, H] tensor")
            self.local = local if local is not None else torch.zeros(self.V, self.H, dtype=torch.float32, device=device)
        self._n, self._k = 1, 0

    @property
    def shard(self) -> torch.Tensor:
        return self.buf.local

    def begin_step(self, n_micro_batches: int = 1) -> None:
        if n_micro_batches < 1:
            raise ValueError("a step has at least one micro-batch")
        self._n, self._k = int(n_micro_batches), 0
        self.buf.local.zero_()
        if self.local is not None and self._n > 1:
            self.local.zero_()
        self.buf.barrier()

    def accumulate(self, hidden, weight, ids, lse, dlogp, inv_temp, shift=True, round_bf16=True, want_dhidden=True, workspace=None):
        """the LM-head backward of one micro-batch with its d_weight going where the schedule says; returns d_hidden"""
        from .lm_head import lm_head_backward_dw_scatter_raw, lm_head_backward_raw

        self._k += 1
        if self._k > self._n:
            raise RuntimeError(f"ShardedGradAccumulator: micro-batch {self._k} of a step announced with {self._n} (call begin_step(n))")
        if self.every_micro_batch or self._k == self._n:
            add = self.local if (self.local is not None and self._n > 1) else None
            return lm_head_backward_dw_scatter_raw(hidden, weight, ids, lse, dlogp, inv_temp, self, shift, round_bf16, want_dhidden=want_dhidden,
                                                   local_add=add, workspace=workspace)
        dh, _ = lm_head_backward_raw(hidden, weight, ids, lse, dlogp, inv_temp, shift, round_bf16, want_dhidden=want_dhidden, dweight=self.local,
                                     dweight_accumulate=True, workspace=workspace)
        return dh

    def skip(self, weight: torch.Tensor) -> None:
        """a micro-batch without any active position (the orchestrator's all-masked padding batches, batch.py:153-161) still counts
        towards the step: a one-token backward with a zero upstream gradient keeps the schedule (and, when it is the last
        micro-batch, sends the local total)"""
        dev = weight.device
        self.accumulate(torch.zeros(1, self.H, dtype=torch.bfloat16, device=dev), weight, torch.zeros(1, dtype=torch.int64, device=dev),
                        torch.zeros(1, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.float32, device=dev), 1.0, shift=False,
                        want_dhidden=False)

    def finish_step(self) -> torch.Tensor:
        if self._k != self._n:
            raise RuntimeError(f"ShardedGradAccumulator: step finished after {self._k} of {self._n} micro-batches")
        self.buf.barrier()
        return self.buf.local

    def close(self) -> None:
        self.buf.close()
