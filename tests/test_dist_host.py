"""N > 1 host logic on CPU: world_size-2 gloo run of the barcode-hash shard + all-to-all exchange (dropest_b200/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_utils as pu
from dropest_b200 import dist as dgdist
from dropest_b200.synth import SynthSpec, SynthTables, rank_of, read_whitelist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spec(n_total):
    return SynthSpec(n_reads=n_total, n_cells=60, n_genes=100, cb_len=16, umi_len=10, whitelist_parts=read_whitelist(pu.WL_SYNTH_7_9), seed=5)


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = n_total // world
    mine = SynthTables(_spec(n_total)).generate_host(rank * per, per)  # this rank's slice of the global stream
    routed, counts = dgdist.route_host(mine, world)
    t = torch.from_numpy(np.frombuffer(routed.tobytes(), dtype=np.uint8).copy())
    recv, n = dgdist.exchange(t, counts)
    got = dgdist.records_from_tensor(recv)
    assert got.shape[0] == n
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), got)
    dist.barrier()
    dist.destroy_process_group()


def test_barcode_hash_all_to_all_world2(tmp_path):
    world, n_total = 2, 40000
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    whole = SynthTables(_spec(n_total)).generate_host(0, n_total)
    allrecs = np.concatenate(parts)
    assert allrecs.shape[0] == n_total  # nothing lost, nothing duplicated
    np.testing.assert_array_equal(allrecs[np.argsort(allrecs["read_idx"])], whole)
    # every barcode lives wholly on the rank its hash names: per-rank grouping is then exact
    for r, p in enumerate(parts):
        assert np.all(rank_of((p["key"] >> np.uint64(24)).astype(np.uint64), world) == r)
    assert min(len(p) for p in parts) > 0.3 * n_total
