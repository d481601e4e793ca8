"""Multi-GPU plumbing: shard reads by barcode hash, one all-to-all-v, then per-rank independent grouping.

torch.distributed is only the transport here (NCCL on GPUs, gloo in the CPU tests); the partition itself is our kernel
(dge_route_by_barcode_device, csrc/synth.cu) or, on the CPU test path, its numpy mirror below.
"""
from __future__ import annotations

import ctypes as C
import os
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


def route_count_slices(device: int, in_ptr: int, n: int, world: int, slice_len: int, n_slices: int, cursors_ptr: int, stream: int = 0) -> np.ndarray:
    """dge_route_count_slices_device: counts[n_slices, world] on the host, per-slice segment prefixes in the caller's device scratch."""
    counts = np.zeros((n_slices, world), dtype=np.uint64)
    rc = load_library().dge_route_count_slices_device(device, C.c_void_p(in_ptr), n, world, slice_len, n_slices, counts.ctypes.data,
                                                      C.c_void_p(cursors_ptr), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"dge_route_count_slices_device failed with {rc}")
    return counts


def route_scatter_slice(device: int, in_ptr: int, n_slice: int, world: int, cursors_ptr: int, out_ptr: int, stream: int = 0):
    rc = load_library().dge_route_scatter_slice_device(device, C.c_void_p(in_ptr), n_slice, world, C.c_void_p(cursors_ptr), C.c_void_p(out_ptr),
                                                       C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"dge_route_scatter_slice_device failed with {rc}")


class PipelinedExchange:
    """Barcode-hash routing + all-to-all in slices, overlapped with the fill of the slices already received:
        slice s: route kernel (our stream) -> all_to_all_single (NCCL's stream, async) -> dge_add_batch_device (our stream)
    The per-slice split sizes of ALL slices come from one counting pass + one all-gather, so the loop itself never waits for the host."""

    def __init__(self, device: int, n: int, world: int, n_slices: int = 8, slack: float = 1.25, group=None):
        import torch

        self.device, self.n, self.world, self.group = device, n, world, group
        self.slice_len = ((n + n_slices - 1) // n_slices + 2047) // 2048 * 2048
        self.n_slices = (n + self.slice_len - 1) // self.slice_len
        dev = f"cuda:{device}"
        self.routed = torch.empty(n * 16, dtype=torch.uint8, device=dev)
        self.recv_cap = int(self.slice_len * slack) + 4096
        self.recv = torch.empty(self.n_slices * self.recv_cap * 16, dtype=torch.uint8, device=dev)
        self.cursors = torch.empty(self.n_slices * 64, dtype=torch.int64, device=dev)

    def run(self, cont, raw_ptr: int, stream) -> int:
        """Routes, exchanges and fills; returns the number of records this rank owns.  The receive buffers are referenced by the
        container until its set_initialized returns."""
        import torch
        import torch.distributed as dist

        world, S = self.world, self.n_slices
        sp = stream.cuda_stream
        counts = route_count_slices(self.device, raw_ptr, self.n, world, self.slice_len, S, self.cursors.data_ptr(), sp)
        mine = torch.from_numpy(counts.astype(np.int64).reshape(-1)).to(self.routed.device)
        allc = torch.empty(world * mine.numel(), dtype=torch.int64, device=self.routed.device)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        allc = allc.cpu().numpy().reshape(world, S, world)          # [source, slice, destination]; nothing else is queued yet: a short wait
        rank = dist.get_rank(self.group)
        total, pending = 0, None
        for s in range(S):
            n_slice = min(self.slice_len, self.n - s * self.slice_len)
            off = s * self.slice_len * 16
            route_scatter_slice(self.device, raw_ptr + off, n_slice, world, self.cursors.data_ptr() + s * 64 * 8, self.routed.data_ptr() + off, sp)
            in_split = [int(x) * 16 for x in counts[s]]
            out_split = [int(x) * 16 for x in allc[:, s, rank]]
            got = sum(out_split) // 16
            if got > self.recv_cap:
                raise RuntimeError("receive slice buffer too small: raise `slack`")
            src = self.routed[off: off + sum(in_split)]
            dst = self.recv[s * self.recv_cap * 16: s * self.recv_cap * 16 + got * 16]
            # NCCL's stream waits for what is queued on ours so far (= this slice's scatter), not for the later slices
            work = dist.all_to_all_single(dst, src, output_split_sizes=out_split, input_split_sizes=in_split, group=self.group, async_op=True)
            if pending is not None:
                pending[0].wait()                                    # our stream waits for that slice's all-to-all only
                cont.add_batch_device(pending[1], pending[2])
            pending = (work, dst.data_ptr(), got)
            total += got
        pending[0].wait()
        cont.add_batch_device(pending[1], pending[2])
        return total


class PeerExchange:
    """Barcode-hash routing with the exchange fused into the kernels on both sides of it (one process per GPU of one NVLink / NVSwitch node):
        single-pass scatter by owner: most tiles go into per-destination windows of THIS rank's HBM that every peer has mapped (CUDA IPC),
            `push16` of every 16 tiles store their remote records straight into the owner's HBM (NVLink is idle during the scatter otherwise)
        -> all-gather of the segment sizes (doubles as the "scatter done" barrier)
        -> ONE fill launch per owner over its own window, the windows the peers pushed into, and the peers' windows, which its bulk-async
           copies pull over NVLink while local tiles are being processed
    No all-to-all pass, no staging copy: every record crosses NVLink exactly once, inside a kernel that does other work at the same time."""

    def __init__(self, device: int, n: int, world: int, group=None, fused: bool = True, slack: float = 1.25, push16=None):
        import torch
        import torch.distributed as dist

        self.device, self.n, self.world, self.group, self.fused = device, n, world, group, fused
        self.rank = dist.get_rank(group)
        # share of the remote records that travels during the scatter: about what NVLink can carry while the scatter streams 32 B per
        # record through local HBM (measured: 2.6 ms for 400 M records)
        if push16 is None:
            push16 = int(os.environ.get("DGE_PUSH16", round(16 * min(0.5, 0.25 * world / max(1, world - 1))))) if world > 1 else 0
        self.push16 = max(0, min(16, int(push16)))
        # every destination owns a window of `cap` records in the routed buffer; the peers' pushes land in `world` more windows behind them.
        # A skewed batch that overflows a window is routed again with exact counts (dense layout in the first part: n records always fit)
        self.cap = (int(max(n, 1) / world * slack) + 4096 + 15) // 16 * 16
        self.push_cap = self.cap if self.push16 else 0
        lib = load_library()
        handle = C.create_string_buffer(64)
        p = C.c_void_p()
        records = max(self.cap * world, max(n, 1)) + self.push_cap * world
        if lib.dge_peer_alloc(device, records * 16, C.byref(p), handle) != 0:
            raise RuntimeError("dge_peer_alloc failed: " + lib.dge_last_error(None).decode())
        self.routed_ptr = p.value
        self.recv_off = max(self.cap * world, max(n, 1))            # first record of the push windows (one per source rank)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.peer_ptr = []
        for r in range(world):
            if r == self.rank:
                self.peer_ptr.append(self.routed_ptr)
                continue
            q = C.c_void_p()
            if lib.dge_peer_open(device, handles[r], C.byref(q)) != 0:
                raise RuntimeError("dge_peer_open failed: " + lib.dge_last_error(None).decode())
            self.peer_ptr.append(q.value)
        dev = f"cuda:{device}"
        self.cursors = torch.empty(64, dtype=torch.int64, device=dev)
        self.state = torch.empty(2 * world + 1, dtype=torch.int64, device=dev)
        # where this rank's pushes land in every peer: its window (index = source rank) behind the peer's pull windows
        self.push_base = torch.tensor([self.peer_ptr[d] + (self.recv_off + self.rank * self.push_cap) * 16 for d in range(world)], dtype=torch.int64, device=dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bytes_pulled = self.bytes_pushed = 0
        self.n_fallbacks = 0

    def close(self):
        lib = load_library()
        for r, q in enumerate(self.peer_ptr):
            if r != self.rank and q:
                lib.dge_peer_close(self.device, C.c_void_p(q))
        self.peer_ptr = []
        if self.routed_ptr:
            lib.dge_peer_free(self.device, C.c_void_p(self.routed_ptr))
            self.routed_ptr = None

    def scatter(self, raw_ptr: int, stream_ptr: int):
        rc = load_library().dge_route_scatter_bounded_device(self.device, C.c_void_p(raw_ptr), self.n, self.world, self.rank, self.cap,
                                                             C.c_void_p(self.state.data_ptr()), C.c_void_p(self.routed_ptr), self.push16, self.push_cap,
                                                             C.c_void_p(self.push_base.data_ptr()), C.c_void_p(stream_ptr))
        if rc != 0:
            raise RuntimeError("dge_route_scatter_bounded_device failed")

    def run(self, cont, raw_ptr: int, stream) -> int:
        """Routes and fills; returns the number of records this rank owns.  The peers' buffers are read until this rank's
        set_initialized has run; the next run() starts with a barrier, so a source never overwrites records a peer still needs."""
        import torch
        import torch.distributed as dist

        world, rank = self.world, self.rank
        sp = stream.cuda_stream
        dist.all_reduce(self.token, group=self.group)                # every peer is done with the previous step's buffers
        # single-pass routing; sizes + overflow flag of every rank in one all-gather, which completes after every rank's scatter
        # (stream order on each rank): the "scatter done" barrier, for the local windows and for what was pushed into ours
        self.scatter(raw_ptr, sp)
        W = 2 * world + 1
        alls = torch.empty(world * W, dtype=torch.int64, device=self.token.device)
        dist.all_gather_into_tensor(alls, self.state, group=self.group)
        alls = alls.cpu().numpy().reshape(world, W)
        ptrs, cnts = [], []
        if not alls[:, 2 * world].any():
            pull, push = alls[:, :world], alls[:, world:2 * world]   # [source, destination]
            for k in range(world):
                src = (rank + k) % world                              # own window first, then the peers in rotation
                ptrs.append(self.peer_ptr[src] + rank * self.cap * 16)
                cnts.append(int(pull[src, rank]))
                if src != rank and push[src, rank]:
                    ptrs.append(self.routed_ptr + (self.recv_off + src * self.push_cap) * 16)   # what `src` pushed: already in this rank's HBM
                    cnts.append(int(push[src, rank]))
            self.bytes_pulled = int(pull[:, rank].sum() - pull[rank, rank]) * 16
            self.bytes_pushed = int(push[rank, :].sum()) * 16
        else:
            # a window overflowed somewhere (one barcode with a large share of the reads): every rank routes again with exact counts
            self.n_fallbacks += 1
            dist.all_reduce(self.token, group=self.group)
            counts = route_count_slices(self.device, raw_ptr, self.n, world, ((self.n + 2047) // 2048) * 2048 or 2048, 1, self.cursors.data_ptr(), sp)
            route_scatter_slice(self.device, raw_ptr, self.n, world, self.cursors.data_ptr(), self.routed_ptr, sp)
            mine = torch.from_numpy(counts.astype(np.int64).reshape(-1)).to(self.token.device)
            allc_t = torch.empty(world * world, dtype=torch.int64, device=self.token.device)
            dist.all_gather_into_tensor(allc_t, mine, group=self.group)
            allc = allc_t.cpu().numpy().reshape(world, world)
            for k in range(world):
                src = (rank + k) % world
                ptrs.append(self.peer_ptr[src] + int(allc[src, :rank].sum()) * 16)
                cnts.append(int(allc[src, rank]))
            self.bytes_pulled = int(allc[:, rank].sum() - allc[rank, rank]) * 16
            self.bytes_pushed = 0
        total = sum(cnts)
        if self.fused:
            cont.add_batch_segments_device(ptrs, cnts)                # ONE launch: local and remote tiles alternate inside every block
        else:
            for p, c in zip(ptrs, cnts):
                if c:
                    cont.add_batch_device(p, c)
        return total


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


def _empty_u8(n: int, device):
    import torch

    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=device)


def merge_across_ranks(cont, device: str, group=None):
    """Exact whitelist merge for sharded runs: drives the library's state machine (dge_dist_step) and performs
    the collectives it asks for over torch.distributed (NCCL on GPUs).  Every rank works on its own cells only: one all-gather of
    16-byte target summaries, then three small all-to-alls (candidate pairs with the child lists, intersection sizes, commits).
    Call between cont.set_initialized() and cont.merge_and_filter().  Returns the bytes this rank sent per step."""
    import os
    import time

    import torch
    import torch.distributed as dist

    from .capi import DIST_ALLGATHER, DIST_DONE

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    trace = os.environ.get("DGE_TRACE") and rank == 0
    io = cont.dist_io(world, rank)
    keep = []        # receive buffers stay alive until the step that consumes them has returned
    sent = []
    t_prev = time.perf_counter()
    while True:
        # the library runs on its own stream (or the caller's, set_stream) and returns with that stream synchronised; torch's
        # collectives below are ordered after it by the host, and we synchronise torch's stream before handing buffers back
        kind = cont.dist_step(io)
        if trace:
            now = time.perf_counter()
            print(f"[dge] dist: step {io.stage} (library)          {1000 * (now - t_prev):8.3f} ms", flush=True)
            t_prev = now
        if kind == DIST_DONE:
            break
        if kind == DIST_ALLGATHER:
            mine = int(io.send_bytes[0])
            sizes_t = torch.empty(world, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(sizes_t, torch.tensor([mine], dtype=torch.int64, device=device), group=group)
            sizes = [int(x) for x in sizes_t.cpu().tolist()]
            mx = max(max(sizes), 16)
            send = _empty_u8(mx, device)
            if mine:
                send[:mine].copy_(_as_tensor(io.send, mine, device))
            gathered = _empty_u8(mx * world, device)
            dist.all_gather_into_tensor(gathered[: mx * world], send[:mx], group=group)
            recv = torch.cat([gathered[r * mx: r * mx + sizes[r]] for r in range(world)]) if sum(sizes) else _empty_u8(16, device)
            out_sizes = sizes
            sent.append(mine)
        else:
            in_split = [int(io.send_bytes[d]) for d in range(world)]
            send_sizes = torch.tensor(in_split, dtype=torch.int64, device=device)
            recv_sizes = torch.empty_like(send_sizes)
            dist.all_to_all_single(recv_sizes, send_sizes, group=group)
            out_split = [int(x) for x in recv_sizes.cpu().tolist()]
            total_in, total_out = sum(in_split), sum(out_split)
            send = _as_tensor(io.send, total_in, device) if total_in else _empty_u8(16, device)[:0]
            recv = _empty_u8(total_out, device)
            dist.all_to_all_single(recv[:total_out], send[:total_in], output_split_sizes=out_split, input_split_sizes=in_split, group=group)
            out_sizes = out_split
            sent.append(total_in)
        torch.cuda.current_stream().synchronize()
        keep = [recv]
        io.recv = recv.data_ptr()
        for r in range(world):
            io.recv_bytes[r] = out_sizes[r]
        if trace:
            now = time.perf_counter()
            print(f"[dge] dist: collective after step {io.stage}     {1000 * (now - t_prev):8.3f} ms ({sent[-1]} bytes sent)", flush=True)
            t_prev = now
    del keep
    return {"bytes_sent": sent}


def _as_tensor(ptr: int, nbytes: int, device):
    """A uint8 torch view of `nbytes` of library-owned device memory (valid until the next dge_dist_step)."""
    import torch

    class _Mem:
        pass

    m = _Mem()
    m.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(m, device=device)


def merge_across_handles(conts):
    """The same protocol with every 'rank' a handle in THIS process (all on one GPU or on several): the collectives are plain copies.
    Used by the single-GPU tests of the cross-rank merge and as a reference implementation of the transport."""
    import torch

    from .capi import DIST_ALLGATHER, DIST_DONE

    world = len(conts)
    ios = [c.dist_io(world, r) for r, c in enumerate(conts)]
    devs = [torch.device("cuda", c.cfg.device) for c in conts]
    keep = []
    while True:
        kinds = [c.dist_step(io) for c, io in zip(conts, ios)]
        assert len(set(kinds)) == 1, f"handles disagree on the collective: {kinds}"
        if kinds[0] == DIST_DONE:
            break
        new_keep = []
        if kinds[0] == DIST_ALLGATHER:
            pieces = [(_as_tensor(io.send, int(io.send_bytes[0]), devs[r]).clone() if io.send_bytes[0] else _empty_u8(16, devs[r])[:0]) for r, io in enumerate(ios)]
            for r, io in enumerate(ios):
                recv = torch.cat([p.to(devs[r]) for p in pieces]) if sum(p.numel() for p in pieces) else _empty_u8(16, devs[r])[:0]
                recv = torch.cat([recv, _empty_u8(16, devs[r])])  # never a null pointer
                new_keep.append(recv)
                io.recv = recv.data_ptr()
                for s in range(world):
                    io.recv_bytes[s] = pieces[s].numel()
        else:
            blobs = []
            for r, io in enumerate(ios):
                total = sum(int(io.send_bytes[d]) for d in range(world))
                whole = _as_tensor(io.send, total, devs[r]).clone() if total else _empty_u8(16, devs[r])[:0]
                offs = np.concatenate([[0], np.cumsum([int(io.send_bytes[d]) for d in range(world)])]).astype(np.int64)
                blobs.append([whole[int(offs[d]): int(offs[d + 1])] for d in range(world)])
            for r, io in enumerate(ios):
                parts = [blobs[s][r].to(devs[r]) for s in range(world)]
                recv = torch.cat(parts + [_empty_u8(16, devs[r])])
                new_keep.append(recv)
                io.recv = recv.data_ptr()
                for s in range(world):
                    io.recv_bytes[s] = parts[s].numel()
        torch.cuda.synchronize()
        keep = new_keep
    del keep
