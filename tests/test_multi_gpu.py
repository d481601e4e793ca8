"""2-GPU run (NCCL): barcode-hash shard + all-to-all + per-rank grouping must reproduce the single-GPU count matrix
(merge = none: every barcode lives wholly on one rank, so per-rank results are exact).  Skipped with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest

import parity_utils as pu

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spec(n_total, umi_len=10, n_genes=150):
    from dropest_b200.synth import SynthSpec, read_whitelist

    return SynthSpec(n_reads=n_total, n_cells=80, n_genes=n_genes, cb_len=16, umi_len=umi_len, whitelist_parts=read_whitelist(pu.WL_SYNTH_7_9), seed=9)


def _real_config(dg, umi_len, n_genes, directional, **kw):
    return dg.Config(cb_len=16, umi_len=umi_len, n_genes=n_genes, merge_type=dg.MERGE_REAL, barcodes_type=dg.BARCODES_CONST,
                     barcodes_file=pu.WL_SYNTH_7_9, min_genes_before_merge=5, min_genes_after_merge=10, max_barcodes_hint=1 << 16,
                     umi_merge_type=dg.UMI_MERGE_DIRECTIONAL if directional else dg.UMI_MERGE_SIMPLE, **kw)


def _triplets(c, dg):
    cells = c.cells(dg.CELLS_REAL)
    indptr, genes, vals = c.matrix(dg.MATRIX_CM_RAW)
    col = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
    return np.stack([cells["barcode"][col].astype(np.uint64), genes.astype(np.uint64), vals.astype(np.uint64)], axis=1)


def _cm_triplets(c, dg, which_cells, which_matrix):
    cells = c.cells(which_cells)
    indptr, genes, vals = c.matrix(which_matrix)
    col = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
    return np.stack([cells["barcode"][col].astype(np.uint64), genes.astype(np.uint64), vals.astype(np.uint64)], axis=1)


def _worker_real(rank, world, port, n_total, out_dir, umi_len, n_genes, directional, exchange):
    import torch
    import torch.distributed as dist

    import dropest_b200 as dg
    from dropest_b200 import dist as dgdist
    from dropest_b200.synth import SynthTables

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    per = n_total // world
    raw = torch.empty(per * 16, dtype=torch.uint8, device=f"cuda:{rank}")
    SynthTables(_spec(n_total, umi_len, n_genes)).generate_device(rank, rank * per, per, raw.data_ptr())
    c = dg.Container(_real_config(dg, umi_len, n_genes, directional, device=rank, sharded=True))
    stream = torch.cuda.current_stream()
    c.set_stream(stream.cuda_stream)
    if exchange.startswith("peer"):   # the fill kernel pulls the routed records out of the other rank's HBM (CUDA IPC peer memory over NVLink)
        # "peer_overflow": destination windows too small on purpose -> the single-pass routing reports it and the exact two-pass routing takes over
        pipe = dgdist.PeerExchange(rank, per, world, slack=0.55 if exchange == "peer_overflow" else 1.25)
    else:                    # routing + sliced NCCL all-to-all overlapped with the fill
        pipe = dgdist.PipelinedExchange(rank, per, world, n_slices=5)
    cnt = pipe.run(c, raw.data_ptr(), stream)
    torch.cuda.synchronize()
    if exchange.startswith("peer"):
        assert pipe.n_fallbacks == (1 if exchange == "peer_overflow" else 0)
    dgdist.sync_umi_first_seen(c, f"cuda:{rank}")
    c.set_initialized()
    stats = dgdist.merge_across_ranks(c, f"cuda:{rank}")
    c.merge_and_filter()
    np.save(os.path.join(out_dir, f"cm{rank}.npy"), _cm_triplets(c, dg, dg.CELLS_FILTERED, dg.MATRIX_CM))
    np.save(os.path.join(out_dir, f"raw{rank}.npy"), _cm_triplets(c, dg, dg.CELLS_REAL, dg.MATRIX_CM_RAW))
    a, b = c.merge_pairs()
    np.save(os.path.join(out_dir, f"pairs{rank}.npy"), np.stack([a, b], axis=1))
    filt = c.cells(dg.CELLS_FILTERED)
    np.save(os.path.join(out_dir, f"filt{rank}.npy"), np.stack([filt["barcode"], filt["umis_stat"].astype(np.uint64), filt["reads_stat"].astype(np.uint64),
                                                             filt["requested_genes_num"].astype(np.uint64)], axis=1))
    s = c.summary()
    np.save(os.path.join(out_dir, f"sum{rank}.npy"), np.array([s[k] for k in ("n_merged", "n_excluded", "n_unresolved", "real_cells_number", "filtered_cells_number",
                                                                                "n_umis_merged")]))
    c.close()
    torch.cuda.synchronize()
    dist.barrier()
    if exchange.startswith("peer"):
        pipe.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("umi_len,n_genes,directional,exchange", [(10, 150, False, "peer"), (10, 150, False, "peer_overflow"), (10, 150, False, "nccl"), (5, 20, True, "peer")])
def test_two_gpu_cross_rank_whitelist_merge_is_exact(tmp_path, umi_len, n_genes, directional, exchange):
    """Sharded run + cross-rank merge (dge_dist_*): the union of the per-rank results equals the single-GPU result.
    With the directional UMI merge the per-UMI first-seen table is min-reduced across ranks first."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import dropest_b200 as dg
    from dropest_b200.synth import SynthTables

    world, n_total = 2, 300_000
    mp.spawn(_worker_real, args=(world, _free_port(), n_total, str(tmp_path), umi_len, n_genes, directional, exchange), nprocs=world, join=True)
    recs = SynthTables(_spec(n_total, umi_len, n_genes)).generate_host(0, n_total)
    c = dg.Container(_real_config(dg, umi_len, n_genes, directional))
    c.add_batch(recs)
    c.set_initialized()
    c.merge_and_filter()
    order = lambda t: t[np.lexsort(tuple(t[:, k] for k in reversed(range(t.shape[1]))))]
    cat = lambda name: np.concatenate([np.load(tmp_path / f"{name}{r}.npy") for r in range(world)])
    s = c.summary()
    sums = sum(np.load(tmp_path / f"sum{r}.npy") for r in range(world))
    print("sums", sums, "single", s)
    assert s["n_merged"] > 0
    assert list(sums) == [s["n_merged"], s["n_excluded"], 0, s["real_cells_number"], s["filtered_cells_number"], s["n_umis_merged"]]
    assert (s["n_umis_merged"] > 0) == directional
    a, b = c.merge_pairs()
    np.testing.assert_array_equal(order(cat("pairs")), order(np.stack([a, b], axis=1)))
    np.testing.assert_array_equal(order(cat("cm")), order(_cm_triplets(c, dg, dg.CELLS_FILTERED, dg.MATRIX_CM)))
    np.testing.assert_array_equal(order(cat("raw")), order(_cm_triplets(c, dg, dg.CELLS_REAL, dg.MATRIX_CM_RAW)))
    filt = c.cells(dg.CELLS_FILTERED)
    single_f = np.stack([filt["barcode"], filt["umis_stat"].astype(np.uint64), filt["reads_stat"].astype(np.uint64),
                         filt["requested_genes_num"].astype(np.uint64)], axis=1)
    np.testing.assert_array_equal(order(cat("filt")), order(single_f))
    c.close()


def _worker(rank, world, port, n_total, out_dir):
    import torch
    import torch.distributed as dist

    import dropest_b200 as dg
    from dropest_b200 import dist as dgdist
    from dropest_b200.synth import SynthTables

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    per = n_total // world
    raw = torch.empty(per * 16, dtype=torch.uint8, device=f"cuda:{rank}")
    SynthTables(_spec(n_total)).generate_device(rank, rank * per, per, raw.data_ptr())
    routed = torch.empty_like(raw)
    counts = dgdist.route_device(rank, raw.data_ptr(), per, world, routed.data_ptr())
    got, cnt = dgdist.exchange(routed, counts)
    torch.cuda.synchronize()
    c = dg.Container(dg.Config(cb_len=16, umi_len=10, n_genes=150, device=rank, merge_type=dg.MERGE_NONE, min_genes_before_merge=5,
                               min_genes_after_merge=5, sharded=True, max_barcodes_hint=1 << 16))
    c.add_batch_device(got.data_ptr(), cnt, keepalive=got)
    c.set_initialized()
    c.merge_and_filter()
    np.save(os.path.join(out_dir, f"trip{rank}.npy"), _triplets(c, dg))
    np.save(os.path.join(out_dir, f"sum{rank}.npy"), np.array([c.summary()[k] for k in ("total_cells_number", "real_cells_number", "intergenic_reads", "n_umigs")]))
    c.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_reproduce_single_gpu_matrix(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import dropest_b200 as dg
    from dropest_b200.synth import SynthTables

    world, n_total = 2, 200_000
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    recs = SynthTables(_spec(n_total)).generate_host(0, n_total)
    c = dg.Container(dg.Config(cb_len=16, umi_len=10, n_genes=150, merge_type=dg.MERGE_NONE, min_genes_before_merge=5, min_genes_after_merge=5,
                               max_barcodes_hint=1 << 16))
    c.add_batch(recs)
    c.set_initialized()
    c.merge_and_filter()
    single = _triplets(c, dg)
    s = c.summary()
    c.close()
    multi = np.concatenate([np.load(tmp_path / f"trip{r}.npy") for r in range(world)])
    order = lambda t: t[np.lexsort((t[:, 1], t[:, 0]))]
    sums = sum(np.load(tmp_path / f"sum{r}.npy") for r in range(world))
    print("per-rank sums", [np.load(tmp_path / f"sum{r}.npy").tolist() for r in range(world)], "single", s)
    np.testing.assert_array_equal(order(multi), order(single))
    assert list(sums) == [s["total_cells_number"], s["real_cells_number"], s["intergenic_reads"], s["n_umigs"]]
