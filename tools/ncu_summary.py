"""Summarise an .ncu-rep (pulled back from the GPU box) into the few numbers the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_ncu_x.txt [profiles/roofline_traffic.json key]

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps, per captured launch: duration, DRAM bytes
read/written, DRAM throughput %, tensor-pipe active %, registers, grid.  With a JSON path + key, also stores the
mean DRAM traffic per launch under that key (bench.py reports it as roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "launch__registers_per_thread": "regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__cycles_elapsed.max": "cycles",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines, traffic = [], []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        d = {}
        for k, short in WANT.items():
            if k in col and r[col[k]] != "":
                try:
                    v = float(r[col[k]].replace(",", ""))
                except ValueError:
                    continue
                d[short] = v * UNIT_SCALE.get(units[col[k]], 1)
        tr = d.get("dram_read", 0) + d.get("dram_write", 0)
        traffic.append(tr)
        lines.append(f"{name[:90]}\n    duration {d.get('duration', 0):9.1f} us | DRAM read {d.get('dram_read', 0)/1e6:9.2f} MB "
                     f"write {d.get('dram_write', 0)/1e6:9.2f} MB ({d.get('dram_pct', 0):5.1f} % of peak) | tensor pipe active "
                     f"{d.get('tensor_pipe_active_pct', 0):5.1f} % | SM throughput {d.get('sm_throughput_pct', 0):5.1f} % | "
                     f"regs {d.get('regs', 0):.0f} | cycles {d.get('cycles', 0):.0f}")
    with open(out, "w") as f:
        f.write(f"summary of {rep} (ncu --set full --clock-control none)\n\n" + "\n".join(lines) + "\n")
    print("\n".join(lines))
    if len(sys.argv) > 4 and traffic:
        path, key = sys.argv[3], sys.argv[4]
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
        data[key] = {"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic), "source": out}
        json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
