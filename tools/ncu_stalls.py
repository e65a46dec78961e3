"""Top stalled SASS lines of a captured kernel (`ncu --set full --import-source on`), from `ncu --page source --csv`.

    python tools/ncu_stalls.py gpurun_out/prof_x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    print(rows[hi - 1][1][:120] if hi > 0 and len(rows[hi - 1]) > 1 else "")
    h = rows[hi]
    col = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[hi + 1:] if len(r) > col["# Samples"]]
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[col["# Samples"]] or 0) for r in data)
    sums = {n: sum(int(r[col[n]] or 0) for r in data) for n in stall_cols}
    print(f"total samples {tot}; by reason: " + ", ".join(f"{n[6:]} {100 * v / max(tot, 1):.1f}%" for n, v in sorted(sums.items(), key=lambda kv: -kv[1])[:8]))
    idx = {id(r): i for i, r in enumerate(data)}
    for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:top_n]:
        n = int(r[col["# Samples"]] or 0)
        why = max(stall_cols, key=lambda c: int(r[col[c]] or 0))
        print(f"{n:6d} {100 * n / max(tot, 1):5.1f}%  line {idx[id(r)]:5d}  {why[6:]:14s} {r[col['Source']][:100]}")


if __name__ == "__main__":
    main()
