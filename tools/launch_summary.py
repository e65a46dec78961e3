"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/r01_launch_list_summary.txt ["header line"]

Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolute times."""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([A-Za-z_0-9]+(?:<[^(]*>)?)", name)
    return m.group(1) if m else name[:60]


def main():
    src, out = sys.argv[1], sys.argv[2]
    header = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = {h: i for i, h in enumerate(r)}
            continue
        if len(r) < len(hdr):
            continue
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[hdr["Metric Unit"]]
        v = float(r[hdr["Metric Value"]].replace(",", ""))
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((short(r[hdr["Kernel Name"]]), us))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for n, us in rows:
        tot[n] += us
        cnt[n] += 1
    total = sum(tot.values()) or 1.0
    with open(out, "w") as f:
        if header:
            f.write(header + "\n")
        f.write("(per-launch times are cold-cache and serialised under ncu: compare SHARES)\n\n")
        for n in sorted(tot, key=lambda k: -tot[k]):
            line = (f"{n:46s} launches={cnt[n]:5d} total={tot[n] / 1e3:9.3f} ms share={100 * tot[n] / total:5.1f}% "
                    f"avg={tot[n] / cnt[n]:8.1f} us")
            f.write(line + "\n")
            print(line)
        f.write(f"\n{len(rows)} launches, {total / 1e3:.3f} ms in total\n")


if __name__ == "__main__":
    main()
