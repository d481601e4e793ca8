"""Multi-GPU plumbing: shard reads by barcode hash, one all-to-all-v, then per-rank independent grouping.

torch.distributed is only the transport here (NCCL on GPUs, gloo in the CPU tests); the partition itself is our kernel
(dge_route_by_barcode_device, csrc/synth.cu) or, on the CPU test path, its numpy mirror below.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from .capi import RECORD_DTYPE, load_library
from .synth import rank_of


def route_host(recs: np.ndarray, world: int) -> Tuple[np.ndarray, np.ndarray]:
    """numpy mirror of dge_route_by_barcode_device: records grouped by owner rank (segment order), per-rank counts."""
    owner = rank_of((recs["key"] >> np.uint64(24)).astype(np.uint64), world)
    order = np.argsort(owner, kind="stable")
    return recs[order], np.bincount(owner, minlength=world).astype(np.uint64)


def route_device(device: int, in_ptr: int, n: int, world: int, out_ptr: int, stream: int = 0) -> np.ndarray:
    counts = np.zeros(world, dtype=np.uint64)
    rc = load_library().dge_route_by_barcode_device(device, C.c_void_p(in_ptr), n, world, C.c_void_p(out_ptr), counts.ctypes.data,
                                                    C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"dge_route_by_barcode_device failed with {rc}")
    return counts


def exchange(routed, counts: np.ndarray, recv=None, group=None):
    """ONE all-to-all-v of 16-byte records.  `routed`: uint8 tensor (n*16) already grouped by destination rank, `counts`:
    records per destination.  Returns (uint8 tensor view of the received records, number of records)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    send = torch.tensor(counts.astype(np.int64), device=routed.device)
    got = torch.empty_like(send)
    dist.all_to_all_single(got, send, group=group)
    in_split = [int(x) * 16 for x in counts]
    out_split = [int(x) * 16 for x in got.cpu().tolist()]
    total = sum(out_split)
    if recv is None:
        recv = torch.empty(total, dtype=torch.uint8, device=routed.device)
    assert total <= recv.numel(), "receive buffer too small"
    assert len(in_split) == world
    dist.all_to_all_single(recv[:total], routed[: sum(in_split)], output_split_sizes=out_split, input_split_sizes=in_split, group=group)
    return recv[:total], total // 16


def records_from_tensor(t) -> np.ndarray:
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=RECORD_DTYPE)
