#!/usr/bin/env python
"""bench.py -- reads/s of the dropEst count-matrix hot path on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" = one full pass of the hot path over one synthetic read stream:
    dge_reset -> dge_add_batch_device (barcode table + key packing) -> dge_set_initialized (grouping, per-cell tables,
    real/filtered cells) -> dge_merge_and_filter (whitelist CB merge, final filter, cm / cm_raw in HBM)
`value`  : device-timed (CUDA events on the launching stream), 16-byte records already resident in HBM.
`e2e`    : the same pass through the C ABI with HOST (pinned) records: H2D inside the timed region + D2H of the count matrix.
`--impl reference` : the reference's own CPU implementation of the path (oracle/_ref = its unmodified sources, else our
                      CPU restatement) on a bounded sample of the same workload.
Workload (config.workload): BASELINE.json configs[1] "10x v3: 400M reads / 10k cells / 30k genes, 16bp CB + 12bp UMI, 1 GPU",
per GPU (weak scaling for N > 1: every rank owns its own 10k-cell population, reads routed by barcode hash with one
NCCL all-to-all inside the timed step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default)
    "c2": {"name": "10x v3: 400M reads / 10k cells / 30k genes, 16bp CB + 12bp UMI, 1 GPU",
           "n_reads": 400_000_000, "n_cells": 10_000, "n_genes": 30_000, "cb_len": 16, "umi_len": 12,
           "min_genes_before": 20, "min_genes_after": 100, "min_frac": 0.2, "max_cb_ed": 2, "merge": "real", "whitelist": "product_7x9",
           "barcodes_type": "const", "max_merge_prob": 1e-4, "max_real_merge_prob": 1e-7,
           "merge_desc": "RealBarcodesMergeStrategy, synthetic product whitelist 2048x3328 (7+9 bp)"},
    # configs[2]: inDrop v3 on the reference's own whitelist (configs/indrop_v3.xml)
    "c3": {"name": "inDrop v3: 200M reads / 5k cells, two-part barcode + whitelist merge, 1 GPU",
           "n_reads": 200_000_000, "n_cells": 5_000, "n_genes": 30_000, "cb_len": 16, "umi_len": 6,
           "min_genes_before": 20, "min_genes_after": 100, "min_frac": 0.2, "max_cb_ed": 2, "merge": "real", "whitelist": "indrop_v3",
           "barcodes_type": "indrop", "max_merge_prob": 1e-5, "max_real_merge_prob": 1e-7,
           "merge_desc": "RealBarcodesMergeStrategy, data/barcodes/indrop_v3 (384 x 384, 8+8 bp), barcodes_type=indrop"},
    # configs[4]: Drop-seq, the per-GPU share of "1B reads / 50k cells on 4 GPUs" (configs/drop_seq.xml: no whitelist; -M = PoissonSimpleMergeStrategy
    # with Tools::CollisionsAdjuster, probabilities 1e-5 / 1e-7)
    "c5": {"name": "Drop-seq: 1B reads / 50k cells, 12bp CB + 8bp UMI with CollisionsAdjuster, 4 GPUs -- per-GPU share 250M reads / 12.5k cells, 1 GPU",
           "n_reads": 250_000_000, "n_cells": 12_500, "n_genes": 30_000, "cb_len": 12, "umi_len": 8,
           "min_genes_before": 20, "min_genes_after": 100, "min_frac": 0.2, "max_cb_ed": 2, "merge": "poisson_simple", "whitelist": None,
           "barcodes_type": "const", "max_merge_prob": 1e-5, "max_real_merge_prob": 1e-7,
           "merge_desc": "PoissonSimpleMergeStrategy (-M) + Tools::CollisionsAdjuster, no whitelist"},
}
WORKLOAD = dict(WORKLOADS["c2"])
ALGO_BYTES_PER_READ = 48  # 3 x 16-byte record (read once, scatter once, re-read once)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two heaviest kernels at the full-size workload, from the
    `ncu --set full` captures summarised in profiles/ (file written by tools/traffic_from_ncu.py); {} when absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return {k: v["dram_bytes_per_launch"] for k, v in json.load(f).items() if isinstance(v, dict)}
    except Exception:
        return {}


from dropest_b200.synth import product_whitelist  # noqa: E402  (the bench whitelist; shared with tests/test_bench_shape.py)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = max(smax, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # "under load" = samples in the upper half of what was seen
        load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_spec(n_reads, n_cells, wl_parts, seed=42):
    from dropest_b200.synth import SynthSpec

    return SynthSpec(n_reads=n_reads, n_cells=n_cells, n_genes=WORKLOAD["n_genes"], cb_len=WORKLOAD["cb_len"], umi_len=WORKLOAD["umi_len"],
                     seed=seed, whitelist_parts=wl_parts)


def sample_case(wl_path, wl_parts, sample_reads):
    """The bounded sample both the CPU reference and (for `parity_on_sample`) the CUDA path run on: a scaled replica of the
    workload -- same reads per cell, same genes, same whitelist, same thresholds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_utils as pu

    n_cells = max(10, int(round(sample_reads * WORKLOAD["n_cells"] / WORKLOAD["n_reads"])))
    spec = make_spec(sample_reads, n_cells, wl_parts, seed=43)
    return pu.Case(name="bench_sample", spec=spec, cb_len=spec.cb_len, umi_len=spec.umi_len, n_genes=spec.n_genes, merge=WORKLOAD["merge"], barcodes=wl_path,
                   barcodes_type=WORKLOAD["barcodes_type"], min_genes_before=WORKLOAD["min_genes_before"], min_genes_after=WORKLOAD["min_genes_after"],
                   max_cb_ed=WORKLOAD["max_cb_ed"], min_frac=WORKLOAD["min_frac"], max_merge_prob=WORKLOAD["max_merge_prob"],
                   max_real_merge_prob=WORKLOAD["max_real_merge_prob"], dump_umis=False, n_batches=3, extra={"max_barcodes_hint": 0})


def cpu_reference_run(wl_path, wl_parts, sample_reads, timeout=1800, keep=False):
    """One timed run of the CPU implementation on the sample (oracle/_ref = the compiled unmodified reference when present)."""
    case = sample_case(wl_path, wl_parts, sample_reads)
    import oracle_io
    from dropest_b200.synth import SynthTables, write_packed

    spec = case.spec
    recs = SynthTables(spec).generate_host(0, sample_reads)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.bin")
        write_packed(path, recs, spec.cb_len, spec.umi_len, spec.n_genes)
        res = oracle_io.run_oracle(path, merge=WORKLOAD["merge"], barcodes=wl_path, barcodes_type=WORKLOAD["barcodes_type"],
                                   min_genes_before=WORKLOAD["min_genes_before"], min_genes_after=WORKLOAD["min_genes_after"], max_cb_ed=WORKLOAD["max_cb_ed"],
                                   min_frac=WORKLOAD["min_frac"], max_merge_prob=WORKLOAD["max_merge_prob"], max_real_merge_prob=WORKLOAD["max_real_merge_prob"],
                                   dump_umis=False, timeout=timeout)
    secs = float(res["t_fill_s"][0] + res["t_init_s"][0] + res["t_merge_s"][0])
    out = {"value": sample_reads / secs, "seconds": secs, "kind": res["_kind"], "cores": 1, "n_cells": spec.n_cells,
           "sample": f"{sample_reads} reads / {spec.n_cells} cells / {spec.n_genes} genes scaled replica of the workload "
                     f"(add_record loop {float(res['t_fill_s'][0]):.2f} s + set_initialized {float(res['t_init_s'][0]):.2f} s + "
                     f"merge_and_filter {float(res['t_merge_s'][0]):.2f} s; single thread: the reference dropest is single-threaded)"}
    if keep:
        out["_case"], out["_recs"], out["_oracle"] = case, recs, res
    return out


def parity_on_sample(cb):
    """The CUDA path on the very sample the CPU reference just ran on, compared field by field (cells in first-seen order, flags,
    merge_targets, per-cell stats, filtered order, cm and cm_raw triplets): tests/parity_utils.assert_parity.  The oracle is only the
    checker here."""
    import parity_utils as pu

    try:
        gpu = pu.gpu_run(cb["_case"], cb["_recs"])
        pu.assert_parity({"case": cb["_case"], "oracle": cb["_oracle"], "gpu": gpu}, check_umigs=False)
        s = gpu["summary"]
        return True, {"n_merged": s["n_merged"], "n_excluded": s["n_excluded"], "cm_nnz": s["cm_nnz"], "filtered_cells_number": s["filtered_cells_number"]}
    except AssertionError as e:
        return False, {"mismatch": str(e)[:300]}


def exchange_diagnostics(pipe, cont, raw, stream, torch, dist, n, world):
    """One-off measurements outside the timed region that name the limiter of the N > 1 step: the routing kernels alone, one
    unsliced all-to-all of the routed records alone (NVLink bound: 16 B x (N-1)/N of the reads leave every GPU), and the fill alone."""
    import numpy as np
    from dropest_b200 import dist as dgdist

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize(); dist.barrier()
    if isinstance(pipe, dgdist.PeerExchange):
        # the routing kernel alone (single-pass scatter into the peer-mapped windows), then the fill alone pulling from the peers
        ev[0].record(stream)
        pipe.scatter(raw.data_ptr(), stream.cuda_stream)
        ev[1].record(stream)
        torch.cuda.synchronize(); dist.barrier()
        cont.reset()
        pipe.run(cont, raw.data_ptr(), stream)
        cont.set_initialized()                      # the library reads its fill-kernel events here
        torch.cuda.synchronize(); dist.barrier()
        ms_pull = cont.timings()["ms_fill_kernel"]
        cont.reset()
        return {"diag_ms_route_kernels": ev[0].elapsed_time(ev[1]), "diag_ms_fill_kernels_pulling": ms_pull,
                "diag_pulled_gbytes_per_gpu": pipe.bytes_pulled / 1e9, "diag_pull_gbs_per_gpu": pipe.bytes_pulled / 1e9 / max(ms_pull / 1e3, 1e-9),
                "diag_pushed_gbytes_per_gpu": pipe.bytes_pushed / 1e9, "diag_push16": float(pipe.push16), "diag_routing_fallbacks": float(pipe.n_fallbacks)}
    ev[0].record(stream)
    dgdist.route_count_slices(pipe.device, raw.data_ptr(), n, world, pipe.slice_len, pipe.n_slices, pipe.cursors.data_ptr(), stream.cuda_stream)
    for s in range(pipe.n_slices):
        n_slice = min(pipe.slice_len, n - s * pipe.slice_len)
        dgdist.route_scatter_slice(pipe.device, raw.data_ptr() + s * pipe.slice_len * 16, n_slice, world, pipe.cursors.data_ptr() + s * 64 * 8,
                                   pipe.routed.data_ptr() + s * pipe.slice_len * 16, stream.cuda_stream)
    ev[1].record(stream)
    torch.cuda.synchronize(); dist.barrier()
    # unsliced all-to-all of the same volume (segments of the first slice layout are not contiguous per destination across slices, so
    # send equal splits of the routed buffer: same bytes on the wire)
    per = (n // world) * 16
    recv = pipe.recv[: per * world]
    ev[2].record(stream)
    dist.all_to_all_single(recv, pipe.routed[: per * world])
    ev[3].record(stream)
    torch.cuda.synchronize()
    ms_route = ev[0].elapsed_time(ev[1])
    ms_a2a = ev[2].elapsed_time(ev[3])
    wire = per * (world - 1)
    return {"diag_ms_route_kernels": ms_route, "diag_ms_all_to_all_unsliced": ms_a2a, "diag_a2a_gbytes_out_per_gpu": wire / 1e9,
            "diag_a2a_gbs_per_gpu": wire / 1e9 / (ms_a2a / 1e3)}


def verify_sharded(args, dg, dgdist, torch, wl_path, wl_parts, dev, rank, world, stream):
    """The N-GPU pipeline (routing, all-to-all, per-rank grouping, cross-rank merge) on one global stream of --verify-reads reads must
    give exactly the single-GPU result: union of the shards' cm / cm_raw triplets (barcode, gene, value), merge pairs and filtered-cell
    rows == those of one handle that sees every read (computed on rank 0)."""
    import pickle

    import numpy as np
    import torch.distributed as dist
    from dropest_b200.synth import SynthTables

    V = (args.verify_reads // world) * world
    cells = max(50, int(round(V * WORKLOAD["n_cells"] / WORKLOAD["n_reads"])))
    tables = SynthTables(make_spec(V, cells, wl_parts, seed=47))
    per = V // world
    buf = torch.empty(per * 16, dtype=torch.uint8, device=f"cuda:{dev}")
    tables.generate_device(dev, rank * per, per, buf.data_ptr())

    def config(sharded):
        return dg.Config(cb_len=WORKLOAD["cb_len"], umi_len=WORKLOAD["umi_len"], n_genes=WORKLOAD["n_genes"], device=dev, merge_type=dg.MERGE_REAL,
                         barcodes_type=dg.BARCODES_INDROP if WORKLOAD["barcodes_type"] == "indrop" else dg.BARCODES_CONST, barcodes_file=wl_path, min_genes_before_merge=WORKLOAD["min_genes_before"], min_genes_after_merge=WORKLOAD["min_genes_after"],
                         max_cb_merge_edit_distance=WORKLOAD["max_cb_ed"], min_merge_fraction=WORKLOAD["min_frac"], sharded=sharded)

    def collect(c):
        def trip(cells_kind, mat):
            cl = c.cells(cells_kind)
            indptr, genes, vals = c.matrix(mat)
            col = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
            return np.stack([cl["barcode"][col].astype(np.uint64), genes.astype(np.uint64), vals.astype(np.uint64)], axis=1)

        filt = c.cells(dg.CELLS_FILTERED)
        a, b = c.merge_pairs()
        return {"cm": trip(dg.CELLS_FILTERED, dg.MATRIX_CM), "raw": trip(dg.CELLS_REAL, dg.MATRIX_CM_RAW), "pairs": np.stack([a, b], axis=1),
                "filt": np.stack([filt["barcode"], filt["umis_stat"].astype(np.uint64), filt["reads_stat"].astype(np.uint64),
                                  filt["requested_genes_num"].astype(np.uint64)], axis=1)}

    c = dg.Container(config(True))
    c.set_stream(stream.cuda_stream)
    pipe = dgdist.PeerExchange(dev, per, world) if args.exchange == "peer" else dgdist.PipelinedExchange(dev, per, world, n_slices=4)
    pipe.run(c, buf.data_ptr(), stream)
    c.set_initialized()
    dgdist.merge_across_ranks(c, f"cuda:{dev}")
    c.merge_and_filter()
    mine = collect(c)
    s = c.summary()
    c.close()
    if args.exchange == "peer":
        torch.cuda.synchronize(); dist.barrier()   # nobody reads this rank's routed buffer any more
        pipe.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(pickle.dumps(mine), gathered, dst=0)
    nm = torch.tensor([s["n_merged"], s["n_excluded"]], device=f"cuda:{dev}")
    dist.all_reduce(nm)
    if rank != 0:
        return None
    full = torch.empty(V * 16, dtype=torch.uint8, device=f"cuda:{dev}")
    tables.generate_device(dev, 0, V, full.data_ptr())
    c1 = dg.Container(config(False))
    c1.add_batch_device(full.data_ptr(), V)
    c1.set_initialized()
    c1.merge_and_filter()
    ref = collect(c1)
    s1 = c1.summary()
    c1.close()
    order = lambda t: t[np.lexsort(tuple(t[:, k] for k in reversed(range(t.shape[1]))))] if t.shape[0] else t
    shards = [pickle.loads(g) for g in gathered]
    match = all(np.array_equal(order(np.concatenate([sh[k] for sh in shards])), order(ref[k])) for k in ("cm", "raw", "pairs", "filt"))
    match = match and int(nm[0]) == s1["n_merged"] and int(nm[1]) == s1["n_excluded"]
    return {"reads": V, "cells": cells, "match": bool(match), "n_merged": s1["n_merged"], "n_excluded": s1["n_excluded"], "cm_nnz": s1["cm_nnz"],
            "what": "union of the %d shards' cm / cm_raw triplets, merge pairs and filtered-cell rows == single-GPU run on the same stream" % world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE configs[1] (default, the headline); c3 / c5 = configs[2] / [4]")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the chosen BASELINE config)")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--cpu-sample-reads", type=int, default=10_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--slices", type=int, default=8, help="N > 1, --exchange nccl: slices of the pipelined route / all-to-all / fill")
    ap.add_argument("--chr", type=int, default=0, metavar="N_CHR",
                    help="N = 1: also feed Stats' per-chromosome counters (a 1-byte chromosome id per read, N_CHR chromosomes; dge_add_batch_chr_device)")
    ap.add_argument("--no-chr-probe", action="store_true", help="skip the two extra steps that measure the optional per-chromosome counters")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: 'peer' = the owners' fill kernels pull the routed records out of the sources' HBM over NVLink (no all-to-all pass); "
                         "'nccl' = scatter -> NCCL all-to-all -> fill in slices")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the sharded == single-GPU check")
    ap.add_argument("--verify-reads", type=int, default=16_000_000)
    args = ap.parse_args()
    WORKLOAD.clear(); WORKLOAD.update(WORKLOADS[args.config])
    args.reads = args.reads or WORKLOAD["n_reads"]
    args.cells = args.cells or WORKLOAD["n_cells"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    hbm_peak, peak_src, _ = measured_peaks()

    cache = os.path.join(tempfile.gettempdir(), f"dge_bench_{os.getuid()}")
    os.makedirs(cache, exist_ok=True)
    from dropest_b200.synth import SynthTables, read_whitelist

    wl_path, wl_parts = None, None
    if WORKLOAD["whitelist"] == "product_7x9":
        wl_path = os.path.join(cache, f"wl_7x9_2048x3328_r{rank}.txt")
        product_whitelist(wl_path)
        wl_parts = read_whitelist(wl_path)
    elif WORKLOAD["whitelist"] == "indrop_v3":
        wl_path = os.path.join(ROOT, "tests", "golden", "ref_barcodes", "indrop_v3")  # copy of the reference's data/barcodes/indrop_v3
        wl_parts = read_whitelist(wl_path, indrop=True)
    config = {"workload": WORKLOAD["name"], "reads_per_gpu": args.reads, "cells_per_gpu": args.cells, "genes": WORKLOAD["n_genes"],
              "cb_len": WORKLOAD["cb_len"], "umi_len": WORKLOAD["umi_len"], "merge": WORKLOAD["merge_desc"],
              "min_genes_before_merge": WORKLOAD["min_genes_before"], "min_genes_after_merge": WORKLOAD["min_genes_after"],
              "l2_policy": "inputs (%.1f GB) larger than L2" % (args.reads * 16 / 1e9),
              "partition": ("barcode-hash; " + ("owners' fill kernels pull the routed records from the sources' HBM over NVLink (CUDA IPC peer memory)" if args.exchange == "peer"
                                              else "sliced NCCL all-to-all overlapped with the fill")) if world > 1 else "single GPU"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        # Every step = one full run of the reference's CPU path on the sample.  The sample is sized so that the per-cell merge cost
        # is represented (>= 250 cells); the number of runs is bounded by a wall-clock budget so the arm ends within a few minutes
        # (a single-threaded CPU run has no warm-up effect beyond paging the binary in: one untimed run, then up to `steps` timed).
        budget_s = float(os.environ.get("DGE_REF_BUDGET_S", "300"))
        t_begin = time.perf_counter()
        vals, last = [], None
        if args.warmup > 0:
            cpu_reference_run(wl_path, wl_parts, min(args.cpu_sample_reads, 1_000_000))
        for it in range(args.steps):
            last = cpu_reference_run(wl_path, wl_parts, args.cpu_sample_reads)
            vals.append(last["value"])
            if time.perf_counter() - t_begin + last["seconds"] * 1.3 > budget_s and len(vals) >= 2:
                break
        v = float(np.mean(vals))
        ref_config = dict(config)
        ref_config["sample_reads"] = args.cpu_sample_reads
        ref_config["sample_cells"] = last["n_cells"]
        ref_config["note"] = ("reference arm: every step runs the CPU path on a %d-read / %d-cell scaled replica of the workload named above "
                              "(same reads per cell, genes, whitelist, thresholds); reads/s is per sample read" % (args.cpu_sample_reads, last["n_cells"]))
        line = {"impl": "reference", "metric": "reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
                "steps_run": len(vals), "warmup": args.warmup, "ms_per_step": 1000.0 * args.cpu_sample_reads / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": ref_config,
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------ our arm
    import torch
    import dropest_b200 as dg

    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    n = args.reads
    # every rank draws its slice of one global stream over world*cells cells (different seed per world size keeps N=1 == configs[1])
    spec = make_spec(n * world, args.cells * world, wl_parts)
    tables = SynthTables(spec)
    raw = torch.empty(n * 16, dtype=torch.uint8, device=f"cuda:{dev}")
    tables.generate_device(dev, rank * n, n, raw.data_ptr())
    torch.cuda.synchronize()

    merge_types = {"real": dg.MERGE_REAL, "simple": dg.MERGE_SIMPLE, "poisson_simple": dg.MERGE_POISSON_SIMPLE, "poisson_real": dg.MERGE_POISSON_REAL}
    if world > 1 and WORKLOAD["merge"] != "real":
        raise SystemExit("only the whitelist merge (config c2 / c3) runs sharded; use --gpus 1 for --config " + args.config)
    cfg = dg.Config(cb_len=WORKLOAD["cb_len"], umi_len=WORKLOAD["umi_len"], n_genes=WORKLOAD["n_genes"], device=dev, merge_type=merge_types[WORKLOAD["merge"]],
                    barcodes_type=dg.BARCODES_INDROP if WORKLOAD["barcodes_type"] == "indrop" else dg.BARCODES_CONST,
                    barcodes_file=wl_path, min_genes_before_merge=WORKLOAD["min_genes_before"], min_genes_after_merge=WORKLOAD["min_genes_after"],
                    max_cb_merge_edit_distance=WORKLOAD["max_cb_ed"], min_merge_fraction=WORKLOAD["min_frac"], max_merge_prob=WORKLOAD["max_merge_prob"],
                    max_real_merge_prob=WORKLOAD["max_real_merge_prob"], sharded=world > 1)
    cont = dg.Container(cfg)
    stream = torch.cuda.current_stream()
    cont.set_stream(stream.cuda_stream)
    lib = dg.load_library()

    pipe = None
    if world > 1:
        from dropest_b200 import dist as dgdist

        pipe = dgdist.PeerExchange(dev, n, world) if args.exchange == "peer" else dgdist.PipelinedExchange(dev, n, world, n_slices=args.slices)
    phase_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(max(1, args.steps))]
    phase_ms = {"ms_route_a2a_fill": 0.0, "ms_group_init": 0.0, "ms_dist_merge": 0.0, "ms_filter_matrices": 0.0}

    owned_reads = n
    chr_ids = None

    def make_chr_ids(n_chr):
        """a chromosome per read (genes live on one chromosome each, reads without a gene fall anywhere), made on the device outside the timed region"""
        ids = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}")
        rec64 = raw.view(torch.int64).view(-1, 2)
        for a in range(0, n, 1 << 26):
            w = rec64[a:a + (1 << 26), 1]
            gene = w & 0xFFFFFF
            idx = (w >> 32) & 0xFFFFFFFF
            ids[a:a + (1 << 26)] = torch.where(gene == 0xFFFFFF, (idx * 40503 >> 3) % n_chr, (gene * 2654435761 >> 13) % max(1, n_chr - 1)).to(torch.uint8)
        return ids

    if args.chr and world == 1:
        chr_ids = make_chr_ids(args.chr)
        config["chromosomes"] = args.chr

    def step(k=None):
        """k = index of the timed step (records the phase events), None = warm-up"""
        ev = phase_ev[k] if k is not None else None
        cont.reset()
        if ev: ev[0].record(stream)
        if world > 1:
            # barcode-hash routing (our kernels); --exchange peer: the fill pulls the routed records out of the sources' HBM over NVLink,
            # --exchange nccl: all-to-all-v over NCCL in slices, overlapped with the fill of the received slices
            cnt = pipe.run(cont, raw.data_ptr(), stream)
        else:
            cnt = n
            if chr_ids is not None:
                cont.add_batch_chr_device(raw.data_ptr(), chr_ids.data_ptr(), n)
            else:
                cont.add_batch_device(raw.data_ptr(), n)
        nonlocal owned_reads
        owned_reads = cnt
        if ev: ev[1].record(stream)
        cont.set_initialized()
        if ev: ev[2].record(stream)
        if world > 1:
            dgdist.merge_across_ranks(cont, f"cuda:{dev}")  # exact cross-rank whitelist merge: every rank works on its own cells
        if ev: ev[3].record(stream)
        cont.merge_and_filter()
        if ev: ev[4].record(stream)
        return cnt

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dedup_ms, dedup_launches, launches, stage = 0.0, 0, 0, {"ms_fill": 0.0, "ms_init": 0.0, "ms_merge": 0.0, "ms_finish": 0.0}
    fill_ms, fill_launches = 0.0, 0
    ev0.record(stream)
    for k in range(args.steps):
        step(k)
        t = cont.timings()
        dedup_ms += t["ms_dedup_kernel"]; dedup_launches += t["n_dedup_launches"]; launches += t["n_kernel_launches"]
        fill_ms += t["ms_fill_kernel"]; fill_launches += t["n_fill_launches"]
        for k in stage:
            stage[k] += t[k]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    for evs in phase_ev[: args.steps]:
        for name, a, b in zip(phase_ms, evs[:-1], evs[1:]):
            phase_ms[name] += a.elapsed_time(b) / args.steps
    clocks = sampler.stop()
    summary = cont.summary()
    if world > 1:
        import torch.distributed as dist

        tmax = torch.tensor([ms], device=f"cuda:{dev}")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    total_reads = n * world * args.steps
    value = total_reads / (ms / 1000.0)

    # ---- the optional per-chromosome Stats counters (1-byte chromosome id per read, SURVEY 8a rows a-R / a4): not part of `value` unless --chr
    # is given; their cost is measured here so that the line says what including them would mean
    chr_extra = None
    if world == 1 and not args.chr and not args.no_chr_probe:
        chr_ids = make_chr_ids(25)
        step(); barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(2):
            step()
        c1.record(stream)
        barrier()
        ms_chr = c0.elapsed_time(c1) / 2
        chr_extra = {"included_in_value": False, "chromosomes": 25, "ms_per_step_with_counters": ms_chr, "reads_per_s_with_counters": n / (ms_chr / 1e3),
                     "note": "dge_add_batch_chr_device + k_chr_stats; the kernel runs at the random-L2-sector rate (profiles/r2_chr_stats_notes.txt)"}
        chr_ids = None
        del c0, c1
    elif world == 1 and args.chr:
        chr_extra = {"included_in_value": True, "chromosomes": args.chr}
    # ---- e2e: host (pinned) records -> C ABI -> count matrix back on the host
    e2e = None
    if not args.no_e2e:
        # the caller's page-locked read arrays: key words and gene|mark words (dge_add_batch_soa; read_idx = stream position)
        rec_dev = raw.view(torch.int64).view(-1, 2)
        host_keys = torch.empty(n, dtype=torch.int64, pin_memory=True)
        host_genes = torch.empty(n, dtype=torch.int32, pin_memory=True)
        host_keys.copy_(rec_dev[:, 0])
        host_genes.copy_((rec_dev[:, 1] & 0xFFFFFFFF).to(torch.int32))
        idx0 = int((rec_dev[0, 1] >> 32) & 0xFFFFFFFF)  # the synthetic stream numbers its reads consecutively from this rank's offset
        del rec_dev
        torch.cuda.synchronize()
        d2h = 0
        # result buffers of the caller: page-locked, sized once from the warm-up result (a user would size them from dge_get_matrix)
        # result buffers of the caller: page-locked, (re)sized by the untimed warm-up call below (a user sizes them from dge_get_matrix)
        bufs = {"indptr": None, "genes": None, "vals": None}

        def ensure_bufs(nc, nnz):
            if bufs["indptr"] is None or bufs["indptr"].numel() < nc + 1:
                bufs["indptr"] = torch.empty(int(nc * 1.2) + 1024, dtype=torch.int64, pin_memory=True)
            if bufs["genes"] is None or bufs["genes"].numel() < nnz:
                bufs["genes"] = torch.empty(int(nnz * 1.2) + 1024, dtype=torch.int32, pin_memory=True)
                bufs["vals"] = torch.empty(int(nnz * 1.2) + 1024, dtype=torch.int32, pin_memory=True)

        host_aos = None
        if world > 1:
            # multi-GPU: the rank's own slice of the stream as 16-byte records on the host; they are copied to the device, routed by
            # barcode hash, exchanged (pipelined all-to-all) and filled -- the same path as `value` plus the H2D copy
            host_aos = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
            host_aos.copy_(raw)
            torch.cuda.synchronize()

        def e2e_step():
            nonlocal d2h
            cont.reset()
            if world > 1:
                raw.copy_(host_aos, non_blocking=True)
                pipe.run(cont, raw.data_ptr(), stream)
            else:
                cont.add_batch_soa_ptr(host_keys.data_ptr(), host_genes.data_ptr(), n, idx0)
            cont.set_initialized()
            if world > 1:
                dgdist.merge_across_ranks(cont, f"cuda:{dev}")
            cont.merge_and_filter()
            nc, nnz = cont.matrix_shape(dg.MATRIX_CM)
            ensure_bufs(nc, nnz)  # allocates in the warm-up call only
            cont.matrix_into(dg.MATRIX_CM, bufs["indptr"].data_ptr(), bufs["genes"].data_ptr(), bufs["vals"].data_ptr())
            d2h = (nc + 1) * 4 + nnz * 8
            return nnz

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            nnz_last = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / max(1, min(args.steps, 3))
        checksum = int(bufs["vals"][:nnz_last].sum())  # the matrix is on the host when the timed region ends; summing it is the reader's work
        if world > 1:
            import torch.distributed as dist

            tt = torch.tensor([dt], device=f"cuda:{dev}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": n * world / dt, "unit": "reads/s", "h2d_bytes_per_step": n * (16 if world > 1 else 12) * world, "d2h_bytes_per_step": int(d2h) * world,
               "note": ("per-rank host records (16 B) -> H2D -> barcode-hash routing + exchange (%s) + fill, cross-rank merge included; " % args.exchange if world > 1 else "")
                       + "cm checksum %d" % checksum}
        del host_keys, host_genes

    # ---- N > 1: the sharded pipeline against a single-GPU run on one global stream (outside the timed region)
    verify = None
    if world > 1 and not args.no_verify:
        verify = verify_sharded(args, dg, dgdist, torch, wl_path, wl_parts, dev, rank, world, stream)
    # ---- N > 1: what the step is made of (CUDA events on the launching stream, mean over the timed steps, max over ranks)
    breakdown = None
    if world > 1:
        import torch.distributed as dist

        diag = exchange_diagnostics(pipe, cont, raw, stream, torch, dist, n, world)
        names = list(phase_ms) + list(diag)
        t = torch.tensor([phase_ms[k] for k in phase_ms] + [diag[k] for k in diag], device=f"cuda:{dev}", dtype=torch.float64)
        every = torch.empty(world * t.numel(), device=f"cuda:{dev}", dtype=torch.float64)
        owned = torch.empty(world, device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_gather_into_tensor(every, t)
        dist.all_gather_into_tensor(owned, torch.tensor([float(owned_reads)], device=f"cuda:{dev}", dtype=torch.float64))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        breakdown = {k: float(v) for k, v in zip(names, t.tolist())}
        # per rank (the step is as slow as its slowest shard: barcode-hash sharding leaves a few per cent of imbalance)
        breakdown["per_rank"] = {"reads_owned": [int(x) for x in owned.tolist()],
                                 **{k: [round(float(every[r * t.numel() + i]), 3) for r in range(world)] for i, k in enumerate(list(phase_ms))}}
    if rank != 0:
        return
    # ---- roofline of the dominant kernel, timed live with CUDA events inside the library (on the launching stream).
    # Two launches compete for "dominant": k_fill_pipe (records -> barcode table + packed keys) and the sub-bucket sort+dedup
    # (k_sort_dedup, its size classes are timed as one unit).  Both are reported; "roofline" is the one with the longer launch.
    n_keys = n - summary["intergenic_reads"] if world == 1 else None
    roof, roof_other = None, None
    if n_keys:
        traffic = measured_traffic()
        cands = []
        if fill_launches:
            # algorithmic bytes of one launch: every 16-byte record read once + one packed 8-byte key written per read with a gene
            algo = n * 16 + n_keys * 8
            per = fill_ms / fill_launches
            cands.append((per, {"kernel": "k_fill_pipe", "bound": "hbm", "achieved": algo / (per / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": algo / (per / 1e3) / 1e9 / hbm_peak, "traffic": traffic.get("k_fill_pipe"), "peak_source": peak_src,
                                "ms_per_launch": per, "algorithmic_bytes_per_launch": algo}))
        if dedup_launches:
            # every grouped key read once (8 B) + every distinct (cell,gene,UMI) written once (8 B key + 4 B value)
            algo = n_keys * 8 + summary["n_umigs"] * 12
            per = dedup_ms / dedup_launches
            cands.append((per, {"kernel": "k_sort_dedup", "bound": "hbm", "achieved": algo / (per / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": algo / (per / 1e3) / 1e9 / hbm_peak, "traffic": traffic.get("k_sort_dedup"), "peak_source": peak_src,
                                "ms_per_launch": per, "algorithmic_bytes_per_launch": algo}))
        cands.sort(key=lambda c: -c[0])
        if cands:
            roof = cands[0][1]
        if len(cands) > 1:
            roof_other = cands[1][1]
    path_gbs = value / world * ALGO_BYTES_PER_READ / 1e9
    line = {"metric": "reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "roofline_second": roof_other,
            "roofline_path": {"bound": "hbm", "achieved": path_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": path_gbs / hbm_peak,
                              "bytes_per_read": ALGO_BYTES_PER_READ, "note": "whole hot path per GPU, BASELINE.md definition"},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "result": {k: summary[k] for k in ("total_cells_number", "real_cells_number", "filtered_cells_number", "n_umigs", "cm_nnz", "n_merged", "n_excluded", "n_unresolved")}}
    if world > 1:
        line["config"]["workload"] = ("10x v3 sharded by barcode hash: %d x (%dM reads / %d cells) = %dM reads / %d cells / 30k genes, 16bp CB + 12bp UMI, %d GPUs "
                                      "(BASELINE configs[3] '4B reads / 100k cells on 8 GPUs' at the per-GPU size of configs[1])"
                                      % (world, n // 1_000_000, args.cells, n * world // 1_000_000, args.cells * world, world))
        line["config"]["merge"] += "; exact cross-rank merge (dge_dist_step: all-gather of target summaries + 3 small all-to-alls), result.* are rank 0's shard"
        line["config"]["exchange"] = ("peer: single-pass scatter by owner into per-destination windows of a CUDA-IPC buffer in the source's HBM, a share of the tiles pushed straight into the owners' HBM -> all-gather of segment sizes -> "
                                      "every owner's k_fill_pipe bulk-copies its segments over NVLink (no all-to-all pass)" if args.exchange == "peer"
                                      else "%d slices: route kernel -> NCCL all_to_all_single (async) -> fill, overlapped" % pipe.n_slices)
        line["step_breakdown_ms"] = breakdown
        line["verify"] = verify
    if chr_extra is not None:
        line["chromosome_stats"] = chr_extra
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_run(wl_path, wl_parts, args.cpu_sample_reads, keep=True)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "reads/s", "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]}
            ok, info = parity_on_sample(cb)
            line["parity_on_sample"] = ok
            line["parity_detail"] = info
        except Exception as e:  # the checker is missing: say so, do not hide it
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 1, "kind": "unavailable", "sample": str(e)[:200]}
            line["parity_on_sample"] = None
    print(json.dumps(line))


if __name__ == "__main__":
    main()
    try:  # leave the process group cleanly (all ranks reach this point: rank 0 prints, the others return from main)
        import torch.distributed as _dist

        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
    except Exception:
        pass
