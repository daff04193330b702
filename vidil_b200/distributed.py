"""One process per GPU: the reference's static sharding, plus ONE collective for the result rows.

Reference behaviour (run_visual_tokenization.py:427-463, run_video_CapFilt.py:237-291, utils.py:258-281):
every rank takes the contiguous slice `step = n // world + 1; [rank*step, min(n, rank*step+step))`, writes its
result dict to `tmp/{rank}.json`, `dist.barrier()`, and rank 0 re-reads all files and merges them with
`dict.update` in rank order.  Here the per-rank dicts travel as length-prefixed UTF-8 JSON through
`all_gather` (NCCL over NVLink on GPUs, gloo in CPU tests) and are merged in the same rank order, so the
merged dict — and the `json.dump(..., indent=4)` file rank 0 writes — is identical.  There is no collective
anywhere else on the path: frames of different videos never interact.
"""
from __future__ import annotations

import datetime
import json
import os

import torch
import torch.distributed as dist

from . import jsonio


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def init_distributed_mode(backend: str | None = None) -> dict:
    """Same environment contract as utils.py:258-281 (RANK / WORLD_SIZE / LOCAL_RANK from torch.distributed.run;
    single-process when absent).  Returns {'rank', 'world_size', 'gpu', 'distributed'}."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        gpu = int(os.environ.get("LOCAL_RANK", 0))
    else:
        return dict(rank=0, world_size=1, gpu=0, distributed=False)
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(gpu)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if not dist.is_initialized():
        kwargs = dict(backend=backend, world_size=world, rank=rank, timeout=datetime.timedelta(seconds=7200))
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", gpu)
        dist.init_process_group(**kwargs)
        dist.barrier()
    return dict(rank=rank, world_size=world, gpu=gpu, distributed=True)


def shard_bounds(n_items: int, world_size: int | None = None, rank: int | None = None):
    """The reference partition (run_visual_tokenization.py:429-431): (start, end) of this rank's slice."""
    world_size = get_world_size() if world_size is None else world_size
    rank = get_rank() if rank is None else rank
    step = n_items // world_size + 1
    start = rank * step
    end = min(n_items, start + step)
    return start, max(start, end)


def all_gather_json(obj, device: torch.device | str | None = None) -> list:
    """Every rank contributes one JSON-serialisable object; every rank gets the list in rank order.

    Two collectives: an all_gather of the int64 payload lengths and an all_gather of the uint8 payloads
    padded to the longest.  Payloads are KBs to a few MB per rank — latency-bound, one NVSwitch hop.
    """
    if not is_dist_avail_and_initialized():
        return [json.loads(json.dumps(obj))]
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    payload = torch.frombuffer(bytearray(json.dumps(obj).encode("utf-8")), dtype=torch.uint8).to(device)
    n = torch.tensor([payload.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    width = max(max(sizes), 1)
    padded = torch.zeros(width, dtype=torch.uint8, device=device)
    padded[:payload.numel()] = payload
    bufs = [torch.empty(width, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return [json.loads(b[:s].cpu().numpy().tobytes().decode("utf-8")) if s else None for b, s in zip(bufs, sizes)]


def all_gather_rows(t: torch.Tensor) -> torch.Tensor:
    """Every rank contributes a [n_r, D] tensor (n_r may differ, 0 allowed); every rank gets the [sum n_r, D] concatenation
    in rank order.  Two collectives (row counts, rows padded to the longest) — used to build the phrase-bank embeddings
    once across the ranks instead of once per rank (run_visual_tokenization.py:84-96 runs the whole bank on every rank)."""
    if not is_dist_avail_and_initialized():
        return t
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    padded = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[:t.shape[0]] = t
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded.contiguous())
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def merge_rank_dicts(per_rank: list) -> dict:
    """`dict.update` in rank order — run_visual_tokenization.py:453-457."""
    merged = {}
    for d in per_rank:
        if d:
            merged.update(d)
    return merged


def gather_and_write(result: dict, path: str | None, device=None) -> dict | None:
    """Gather every rank's {video_id: row} dict; rank 0 merges, optionally writes `path` with the reference's
    `json.dump(..., indent=4)` and returns the merged dict (other ranks return None)."""
    per_rank = all_gather_json(result, device)
    if not is_main_process():
        return None
    merged = merge_rank_dicts(per_rank)
    if path is not None:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as out:
            jsonio.dump_indent4(merged, out)          # == json.dump(merged, out, indent=4), run_visual_tokenization.py:52
    return merged
