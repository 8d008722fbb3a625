"""Top stall-sampled CUDA source lines of one launch in an .ncu-rep (needs --import-source on).

    python tools/ncu_source_top.py report.ncu-rep LAUNCH_INDEX [N]
"""
import csv
import io
import subprocess
import sys


def main():
    import os
    rep, launch = os.path.abspath(sys.argv[1]), int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True, cwd="/tmp").stdout
    rows = list(csv.reader(io.StringIO(txt)))
    fname, hdr, lines = "", None, []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            si = hdr.index("# Samples")
            stall0 = hdr.index("stall_barrier")
            stall1 = hdr.index("stall_wait") + 1
        elif hdr and len(r) == len(hdr) and r[0].strip().isdigit() and r[2] in ("", "-"):
            n = int(r[si] or 0)
            if n:
                stalls = sorted(((int(r[i] or 0), hdr[i][6:]) for i in range(stall0, stall1)), reverse=True)[:2]
                lines.append((n, fname, int(r[0]), r[1].strip()[:110], stalls))
    tot = sum(l[0] for l in lines)
    print("total samples", tot)
    for n, f, ln, s, st in sorted(lines, reverse=True)[:top]:
        print(f"{100*n/tot:5.1f}%  {f}:{ln:<5d} {s}   {[f'{k}:{v}' for v, k in st if v]}")


if __name__ == "__main__":
    main()
