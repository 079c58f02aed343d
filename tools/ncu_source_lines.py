"""Per-source-line hot spots of one kernel of an ncu report captured with --import-source on.

    python tools/ncu_source_lines.py gpurun_out/final/prof_pics8.ncu-rep k_rle_expand [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Line No" in r and "# Samples" in r)
    h = rows[hdr]
    si, ii, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    # rows with a line number are the per-line aggregates of the SASS rows that follow them
    body = [r for r in rows[hdr + 1:] if len(r) > si and r[0].isdigit() and r[si].isdigit()]
    tot = sum(int(r[si]) for r in body) or 1
    toti = sum(int(r[ii]) for r in body) or 1
    print("total samples", tot, "warp instructions", toti)
    agg = {}
    for r in body:
        for i in stalls:
            if r[i].isdigit():
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print("stall mix:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(body, key=lambda r: -int(r[si]))[:top]:
        best = max(stalls, key=lambda i: int(r[i]) if r[i].isdigit() else 0)
        print(f"{r[0]:>5} smp {100 * int(r[si]) / tot:5.1f}%  ins {100 * int(r[ii]) / toti:5.1f}%  {h[best][6:]:<14} {r[src].strip()[:120]}")


if __name__ == "__main__":
    main()
