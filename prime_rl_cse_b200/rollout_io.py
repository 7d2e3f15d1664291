"""The data format either side of the path: the per-step, per-rank micro-batch files the orchestrator writes and the
trainer reads.  Same layout, names and behaviour as the reference:

    <output_dir>/rollouts/step_<n>/rank_<r>.pt  =  torch.save(list[MicroBatch])        (orchestrator.py:356-363)

written to `rank_<r>.tmp` and renamed so a reader never sees a partial file; read by `DataLoader.get_batch()`
(src/prime_rl/trainer/rl/data.py:57-75) after `wait_for_batch()` polls for the path (utils/utils.py:119-130).
`FakeDataLoader` mirrors rl/data.py:26-54.  On top of the reference's interface, `get_batch_on_device` replaces the per-tensor
`.to("cuda")` calls of train.py:275-279 by ONE pinned staging buffer and ONE host-to-device copy per micro-batch, issued on a
side stream so the upload of micro-batch i+1 overlaps the kernels of micro-batch i."""

from __future__ import annotations

import os
import time
from pathlib import Path
from typing import Iterator

import torch

from .data import MicroBatch, fake_micro_batch
from .segments import response_lengths_host

_TENSOR_KEYS = ("input_ids", "position_ids", "advantages", "logprobs", "loss_mask")


def get_rollout_dir(output_dir: Path | str) -> Path:
    """utils/utils.py:259-260."""
    return Path(output_dir) / "rollouts"


def rollout_path(output_dir: Path | str, step: int, rank: int) -> Path:
    return get_rollout_dir(output_dir) / f"step_{step}" / f"rank_{rank}.pt"


def save_step(output_dir: Path | str, step: int, all_data_ranks_batches: list[list[MicroBatch]]) -> list[Path]:
    """Writer side (orchestrator.py:356-363): one file per data rank, tmp + rename.  Every tensor is made to OWN its storage
    before it is serialised: `prepare_batch` hands out micro-batches that are slice views of one step-wide buffer per field, and
    torch.save writes a view's whole backing storage -- without the clone every rank file would carry every rank's tokens."""
    step_path = get_rollout_dir(output_dir) / f"step_{step}"
    step_path.mkdir(parents=True, exist_ok=True)
    paths = []
    for i, batches in enumerate(all_data_ranks_batches):
        batch_path = step_path / f"rank_{i}.pt"
        tmp_path = batch_path.with_suffix(".tmp")
        torch.save([{k: (_owned_cpu(v) if torch.is_tensor(v) else v) for k, v in mb.items()} for mb in batches], tmp_path)
        tmp_path.rename(batch_path)
        paths.append(batch_path)
    return paths


def _owned_cpu(t: torch.Tensor) -> torch.Tensor:
    """CPU copy whose storage holds exactly the tensor's elements (no-op for a CPU tensor that already owns a compact storage)"""
    c = t.detach().cpu()
    if c.untyped_storage().nbytes() != c.numel() * c.element_size() or not c.is_contiguous():
        c = c.clone(memory_format=torch.contiguous_format)
    return c


def wait_for_path(path: Path, interval: float = 1.0, timeout: float | None = None) -> None:
    """utils/utils.py:119-130; `timeout` (seconds) is an addition: the reference waits forever."""
    waited = 0.0
    while not path.exists():
        if timeout is not None and waited >= timeout:
            raise TimeoutError(f"{path} did not appear within {timeout} s")
        time.sleep(interval)
        waited += interval


def _validate(mb: dict, where: str) -> None:
    T = None
    for k in _TENSOR_KEYS:
        if k not in mb or not torch.is_tensor(mb[k]):
            raise ValueError(f"{where}: micro-batch lacks tensor {k!r}")
        if mb[k].dim() != 2 or (T is not None and tuple(mb[k].shape) != T):
            raise ValueError(f"{where}: {k} has shape {tuple(mb[k].shape)}, expected [batch, seq] equal across fields")
        T = tuple(mb[k].shape)
    if "temperature" not in mb:
        raise ValueError(f"{where}: micro-batch lacks 'temperature'")


class DataLoader:
    """Loads the micro-batches the orchestrator serialised for this rank (rl/data.py:57-75)."""

    def __init__(self, output_dir: Path | str, start_step: int, rank: int | None = None):
        self.rollout_dir = get_rollout_dir(output_dir)
        self.current_step = start_step
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)  # reference: get_world().rank

    def get_rollout_path(self) -> Path:
        return self.rollout_dir / f"step_{self.current_step}" / f"rank_{self.rank}.pt"

    def wait_for_batch(self, interval: float = 1.0, timeout: float | None = None) -> None:
        wait_for_path(self.get_rollout_path(), interval, timeout)

    def get_batch(self) -> list[MicroBatch]:
        path = self.get_rollout_path()
        batches = torch.load(path)
        if not isinstance(batches, list):
            raise ValueError(f"{path}: expected a list of micro-batches, got {type(batches).__name__}")
        for i, mb in enumerate(batches):
            _validate(mb, f"{path}[{i}]")
        self.current_step += 1
        return batches

    def get_batch_on_device(self, device, stream: "torch.cuda.Stream | None" = None) -> Iterator[MicroBatch]:
        """Yields this step's micro-batches as device tensors, uploading one ahead of the consumer."""
        return upload_micro_batches(self.get_batch(), device, stream)


class FakeDataLoader:
    """rl/data.py:26-54: `batch_size // micro_batch_size // world_size` random micro-batches per step."""

    def __init__(self, batch_size: int = 2, micro_batch_size: int = 1, seq_len: int = 128, world_size: int | None = None):
        if batch_size % micro_batch_size != 0 or batch_size < micro_batch_size:
            raise ValueError("batch_size must be a positive multiple of micro_batch_size")  # rl/config.py:38-43
        world = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else int(world_size)
        self.batch_size, self.micro_batch_size, self.seq_len = batch_size, micro_batch_size, seq_len
        self.num_micro_batches = batch_size // micro_batch_size // world

    def wait_for_batch(self) -> None:
        return

    def get_batch(self) -> list[MicroBatch]:
        return [fake_micro_batch(self.seq_len, self.micro_batch_size) for _ in range(self.num_micro_batches)]


# --------------------------------------------------------------------------- one copy per micro-batch
def _layout(T: int) -> tuple[dict[str, tuple[int, int, torch.dtype]], int]:
    """byte offsets of the five [1, T] tensors inside one staging buffer (8-byte fields first: everything stays aligned)"""
    off, lay = 0, {}
    for k, dt, sz in (("input_ids", torch.int64, 8), ("position_ids", torch.int64, 8), ("advantages", torch.float32, 4),
                      ("logprobs", torch.float32, 4), ("loss_mask", torch.bool, 1)):
        lay[k] = (off, T * sz, dt)
        off += (T * sz + 15) // 16 * 16
    return lay, off


def pack_host(mb: MicroBatch, pin: bool = True) -> tuple[torch.Tensor, int, int]:
    """The five per-token tensors of a micro-batch in ONE (pinned) uint8 buffer.  Returns (buffer, batch, seq)."""
    B, S = mb["input_ids"].shape
    lay, nbytes = _layout(B * S)
    buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
    for k, (off, n, dt) in lay.items():
        buf[off : off + n].view(dt).copy_(mb[k].to(dt).reshape(-1))
    return buf, B, S


def unpack_device(buf: torch.Tensor, B: int, S: int, temperature: float) -> MicroBatch:
    """Views (no copies) of a staging buffer that has been copied to the device."""
    lay, _ = _layout(B * S)
    out = {k: buf[off : off + n].view(dt).view(B, S) for k, (off, n, dt) in lay.items()}
    out["temperature"] = temperature
    return out  # type: ignore[return-value]


def upload_micro_batches(batches: list[MicroBatch], device, stream: "torch.cuda.Stream | None" = None) -> Iterator[MicroBatch]:
    """Generator over device micro-batches: micro-batch i+1 is packed and copied (one H2D copy from pinned memory, on
    `stream`) while the caller works on micro-batch i; the current stream waits on the copy's event before use."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("upload_micro_batches targets a CUDA device; there is no CPU path")
    copy_stream = stream if stream is not None else torch.cuda.Stream(device=device)

    def start(mb):
        host, B, S = pack_host(mb)
        with torch.cuda.stream(copy_stream):
            dev = host.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        # host-side facts the micro-step would otherwise fetch from the device with a synchronisation (train.py:291 does O(T)
        # of them): segment lengths of the packed row, and how many positions can receive a gradient
        extra = {}
        if B == 1:
            extra = {"response_lengths": response_lengths_host(mb["position_ids"].cpu()), "n_active": int(mb["loss_mask"].reshape(-1)[1:].sum())}
        return dev, host, ev, B, S, float(mb["temperature"]), extra

    nxt = start(batches[0]) if batches else None
    for i in range(len(batches)):
        dev, host, ev, B, S, temp, extra = nxt
        nxt = start(batches[i + 1]) if i + 1 < len(batches) else None
        cur = torch.cuda.current_stream(device)
        cur.wait_event(ev)
        dev.record_stream(cur)
        out = unpack_device(dev, B, S, temp)
        out.update(extra)
        yield out
        del host  # the pinned buffer may be recycled only after its copy was consumed
