#!/usr/bin/env python
"""Times the routing kernels alone on one GPU (count pass, scatter) for several world sizes: python tools/route_bench.py [reads]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dropest_b200 import dist as dgdist
from dropest_b200.synth import SynthSpec, SynthTables, product_whitelist, read_whitelist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
product_whitelist("/tmp/route_wl.txt")
spec = SynthSpec(n_reads=n, n_cells=10000, n_genes=30000, cb_len=16, umi_len=12, whitelist_parts=read_whitelist("/tmp/route_wl.txt"), seed=42)
raw = torch.empty(n * 16, dtype=torch.uint8, device="cuda:0")
SynthTables(spec).generate_device(0, 0, n, raw.data_ptr())
out = torch.empty(n * 16, dtype=torch.uint8, device="cuda:0")
cur = torch.empty(64, dtype=torch.int64, device="cuda:0")
st = torch.cuda.current_stream()
for world in (2, 4, 8):
    for rep in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record(st)
        dgdist.route_count_slices(0, raw.data_ptr(), n, world, ((n + 2047) // 2048) * 2048, 1, cur.data_ptr(), st.cuda_stream)
        e[1].record(st)
        dgdist.route_scatter_slice(0, raw.data_ptr(), n, world, cur.data_ptr(), out.data_ptr(), st.cuda_stream)
        e[2].record(st)
        torch.cuda.synchronize()
    print(f"world {world}: count {e[0].elapsed_time(e[1]):.3f} ms ({n*16/e[0].elapsed_time(e[1])/1e6:.0f} GB/s), scatter {e[1].elapsed_time(e[2]):.3f} ms ({n*32/e[1].elapsed_time(e[2])/1e6:.0f} GB/s)", flush=True)
