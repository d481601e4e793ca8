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


def sync_umi_first_seen(cont, device: str, group=None):
    """Min-reduce the per-UMI first-seen read index across ranks (the reference's UMI StringIndexer is global).  No-op unless the
    configured strategies depend on UMI id order (directional UMI merge).  Call after the last add_batch on every rank."""
    import torch
    import torch.distributed as dist

    n = cont.umi_first_size()
    if n == 0 or dist.get_world_size(group) == 1:
        return 0
    t = torch.empty(n, dtype=torch.int32, device=device)
    cont.umi_first_export(t.data_ptr())
    t.bitwise_xor_(-2 ** 31)            # u32 order -> signed order, so that MIN works on the int32 view
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    t.bitwise_xor_(-2 ** 31)
    torch.cuda.synchronize()
    cont.umi_first_import(t.data_ptr())
    return n


def merge_across_ranks(cont, device: str, group=None):
    """Exact whitelist merge for sharded runs: two all-gathers around three local library steps.
    Call between cont.set_initialized() and cont.merge_and_filter()."""
    import torch
    import torch.distributed as dist

    from .capi import DIST_RESULT_DTYPE

    import os
    import time

    trace = os.environ.get("DGE_TRACE") and dist.get_rank(group) == 0
    t_prev = [time.perf_counter()]

    def mark(what):
        if trace:
            torch.cuda.synchronize()
            now = time.perf_counter()
            print(f"[dge] dist: {what:<28s} {1000 * (now - t_prev[0]):8.3f} ms", flush=True)
            t_prev[0] = now

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nc, ne = cont.dist_export_children()
    mark("export children")
    mine = torch.tensor([nc, ne], dtype=torch.int64, device=device)
    counts = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.cpu().numpy().reshape(world, 2)
    max_nc, max_ne = int(counts[:, 0].max()), int(counts[:, 1].max())
    ncs, nes = counts[:, 0], counts[:, 1]
    tot_nc, tot_ne = int(ncs.sum()), int(nes.sum())
    if tot_nc == 0:
        cont.dist_apply(np.zeros(0, dtype=DIST_RESULT_DTYPE), world, rank, np.zeros(0, dtype=np.uint32))
        return {"children": 0, "entries": 0}
    send_i = torch.zeros(max(max_nc, 1) * 32, dtype=torch.uint8, device=device)
    send_k = torch.zeros(max(max_ne, 1) * 8, dtype=torch.uint8, device=device)
    send_v = torch.zeros(max(max_ne, 1) * 4, dtype=torch.uint8, device=device)
    cont.dist_copy_children(send_i.data_ptr(), send_k.data_ptr(), send_v.data_ptr(), nc, ne)
    all_i = torch.empty(world * send_i.numel(), dtype=torch.uint8, device=device)
    all_k = torch.empty(world * send_k.numel(), dtype=torch.uint8, device=device)
    all_v = torch.empty(world * send_v.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(all_i, send_i, group=group)   # children summaries
    dist.all_gather_into_tensor(all_k, send_k, group=group)   # their (gene|umi) lists ...
    dist.all_gather_into_tensor(all_v, send_v, group=group)   # ... and values
    mark("all-gather children")
    g_i = torch.cat([all_i[r * send_i.numel(): r * send_i.numel() + int(ncs[r]) * 32] for r in range(world)])
    g_k = torch.cat([all_k[r * send_k.numel(): r * send_k.numel() + int(nes[r]) * 8] for r in range(world)])
    g_v = torch.cat([all_v[r * send_v.numel(): r * send_v.numel() + int(nes[r]) * 4] for r in range(world)])
    mark("concatenate")
    local = cont.dist_eval_children(g_i.data_ptr(), tot_nc, g_k.data_ptr(), g_v.data_ptr(), tot_ne)
    mark("eval children")
    res_t = torch.from_numpy(local.view(np.uint8)).to(device)
    all_r = torch.empty(world * res_t.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(all_r, res_t, group=group)    # per-rank best candidate of every child
    child_rank = np.repeat(np.arange(world, dtype=np.uint32), ncs.astype(np.int64))
    mark("all-gather results")
    torch.cuda.current_stream().synchronize()
    cont.dist_apply_device(all_r.data_ptr(), world, rank, child_rank)   # combination over ranks on the device; g_k / g_v stay alive until here
    mark("apply")
    return {"children": tot_nc, "entries": tot_ne}
