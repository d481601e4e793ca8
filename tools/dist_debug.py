import os, sys, traceback
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch, torch.distributed as dist
import dropest_b200 as dg
from dropest_b200 import dist as dgdist
from dropest_b200.synth import SynthSpec, SynthTables, read_whitelist, rank_of
import parity_utils as pu
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
try:
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    n_total = 200000; per = n_total // world
    spec = SynthSpec(n_reads=n_total, n_cells=80, n_genes=150, cb_len=16, umi_len=10, whitelist_parts=read_whitelist(pu.WL_SYNTH_7_9), seed=9)
    t = SynthTables(spec)
    raw = torch.empty(per * 16, dtype=torch.uint8, device=f"cuda:{rank}")
    t.generate_device(rank, rank * per, per, raw.data_ptr())
    host = t.generate_host(rank * per, per)
    print(rank, "synth ok", np.array_equal(dgdist.records_from_tensor(raw), host), flush=True)
    routed = torch.empty_like(raw)
    counts = dgdist.route_device(rank, raw.data_ptr(), per, world, routed.data_ptr())
    exp = np.bincount(rank_of((host["key"] >> np.uint64(24)).astype(np.uint64), world), minlength=world)
    print(rank, "counts", counts, "expected", exp, flush=True)
    r = dgdist.records_from_tensor(routed)
    print(rank, "routed owners ok", np.array_equal(np.sort(r["read_idx"]), np.sort(host["read_idx"])), flush=True)
    got, cnt = dgdist.exchange(routed, counts)
    torch.cuda.synchronize()
    g = dgdist.records_from_tensor(got)
    print(rank, "received", cnt, "all mine", bool(np.all(rank_of((g["key"] >> np.uint64(24)).astype(np.uint64), world) == rank)), flush=True)
    c = dg.Container(dg.Config(cb_len=16, umi_len=10, n_genes=150, device=rank, merge_type=dg.MERGE_NONE, min_genes_before_merge=5, min_genes_after_merge=5, sharded=True, max_barcodes_hint=1 << 16))
    c.add_batch_device(got.data_ptr(), cnt, keepalive=got)
    c.set_initialized(); c.merge_and_filter()
    print(rank, c.summary(), flush=True)
    dist.barrier()
except Exception:
    traceback.print_exc()
    sys.exit(1)
