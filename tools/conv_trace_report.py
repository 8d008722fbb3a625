"""Summarise gpurun_out/conv_trace.jsonl (PFB_CONV_TRACE): per launch, median per-CTA timeline in microseconds."""
import json
import sys

import numpy as np

GHZ = float(sys.argv[2]) if len(sys.argv) > 2 else 1.8
names = {2: "prologue_done", 3: "prod_start", 4: "prod_last_issue", 5: "mma_first_data", 9: "mma_t0_start", 6: "mma_t0_issued", 10: "mma_t1_start",
         7: "mma_t1_issued", 11: "mma_t2_start", 8: "mma_t2_issued", 12: "epi_t0_accfull", 15: "epi_t0_done", 13: "epi_t1_accfull", 16: "epi_t1_done",
         14: "epi_t2_accfull", 17: "epi_t2_done", 21: "epiB_t0_done", 22: "epiB_t1_done", 23: "epiB_t2_done", 18: "final_sync", 19: "cluster_sync"}
order = [2, 3, 5, 9, 6, 12, 15, 21, 10, 7, 13, 16, 22, 11, 8, 4, 14, 17, 23, 18, 19]
for ln, line in enumerate(open(sys.argv[1])):
    d = json.loads(line)
    t = np.array(d.pop("t"), dtype=np.float64).reshape(d["grid"], 32)
    g0 = t[:, 0]
    span_ns = t[:, 20].max() - g0.min()
    start_skew = (g0 - g0.min())
    rel = (t - t[:, 1:2]) / GHZ / 1e3  # us since CTA entry (clock64)
    lead = np.arange(d["grid"]) % 2 == 0
    print(f"--- launch {ln}: {d}  kernel span {span_ns/1e3:.1f} us; CTA start skew med {np.median(start_skew)/1e3:.2f} max {start_skew.max()/1e3:.2f} us; CTA life med {np.median(t[:,20]-g0)/1e3:.1f} us")
    row = []
    for k in order:
        sel = lead if k in (5, 6, 7, 8, 9, 10, 11) else np.ones_like(lead)
        v = rel[sel, k]
        v = v[t[sel, k] > 0]
        if len(v):
            row.append(f"{names[k]}={np.median(v):.1f}/{v.max():.1f}")
    print("   " + "  ".join(row))
