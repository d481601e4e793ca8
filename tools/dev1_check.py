import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import dropest_b200 as dg
from dropest_b200.synth import SynthSpec, SynthTables, read_whitelist
import parity_utils as pu
spec = SynthSpec(n_reads=200000, n_cells=80, n_genes=150, cb_len=16, umi_len=10, whitelist_parts=read_whitelist(pu.WL_SYNTH_7_9), seed=9)
t = SynthTables(spec)
recs = t.generate_host(0, 200000)
for dev in (0, 1):
    try:
        torch.cuda.set_device(dev)
        buf = torch.empty(200000 * 16, dtype=torch.uint8, device=f"cuda:{dev}")
        t.generate_device(dev, 0, 200000, buf.data_ptr())
        got = np.frombuffer(buf.cpu().numpy().tobytes(), dtype=dg.RECORD_DTYPE)
        print(dev, 'synth equal', np.array_equal(got, recs))
        c = dg.Container(dg.Config(cb_len=16, umi_len=10, n_genes=150, device=dev, merge_type=dg.MERGE_NONE, min_genes_before_merge=5, min_genes_after_merge=5, max_barcodes_hint=1 << 16))
        c.add_batch_device(buf.data_ptr(), 200000)
        c.set_initialized(); c.merge_and_filter()
        print(dev, c.summary())
        c.close()
    except Exception as e:
        print(dev, 'ERROR', repr(e))
