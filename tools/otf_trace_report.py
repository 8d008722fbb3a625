#!/usr/bin/env python
"""Phase timeline of corr_onthefly_umma_kernel from a PFB_OTF_TRACE file (clock64 stamps of every CTA's second work item).

    PFB_OTF_TRACE=gpurun_out/otf_trace.jsonl python tools/time_config4.py ...
    python tools/otf_trace_report.py gpurun_out/otf_trace.jsonl

Slots (csrc/corr_onthefly_umma.cu): 0 item start (epilogue), 1 next item's region published, 2 number of bands, 3 level
written out; per band kb at 8 + 6 kb: +0 the producer starts issuing the band's loads, +1 the issuer has issued the band's
MMAs, +2 accumulator ready (epilogue thread 0), +3 dump done, +4 gather done.  The roles run decoupled, so "tma" / "mma" are
issue-to-issue and issue-to-ready spans (they include queueing); "period" is the time between consecutive bands leaving
the epilogue, the number that bounds the kernel.
"""
import json
import statistics as st
import sys


def main():
    path = sys.argv[1]
    launches = [json.loads(line) for line in open(path) if line.strip()]
    for li, rec in enumerate(launches[:4]):
        g = rec["grid"]
        s = rec["stamps"]
        rows = {k: [] for k in ("region", "tma", "mma", "dump", "gather", "band", "period", "out", "item", "nb")}
        for c in range(g):
            t = s[c * 64:(c + 1) * 64]
            if not t[0] or not t[3]:
                continue
            nb = int(t[2])
            rows["nb"].append(nb)
            rows["region"].append(t[1] - t[0])
            rows["item"].append(t[3] - t[0])
            last = t[1]
            for kb in range(min(nb, 8)):
                b = t[8 + 6 * kb: 8 + 6 * kb + 5]
                if not all(b):
                    continue
                rows["tma"].append(b[1] - b[0])
                rows["mma"].append(b[2] - b[1])
                rows["dump"].append(b[3] - b[2])
                rows["gather"].append(b[4] - b[3])
                rows["band"].append(b[4] - b[0])
                if kb > 0:
                    rows["period"].append(b[4] - last)
                last = b[4]
            rows["out"].append(t[3] - last)
        print(f"launch {li}: grid {g}, {len(rows['item'])} CTAs traced, clk (median / p90)")
        for k in ("nb", "region", "tma", "mma", "dump", "gather", "band", "period", "out", "item"):
            v = sorted(rows[k])
            if v:
                print(f"  {k:7s} {st.median(v):9.0f} {v[int(0.9 * (len(v) - 1))]:9.0f}   n={len(v)}")


if __name__ == "__main__":
    main()
