"""Turn an `ncu --set full` capture into the numbers bench.py and DESIGN.md quote.

  python tools/ncu_traffic.py gpurun_out/r02_prof_gemm_h3.ncu-rep --kernel gemm_h3_kernel --D 4096 --B 4096 \
      --summary profiles/r02_ncu_gemm_h3_summary.txt

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU), prints per-launch duration, DRAM bytes, tensor-pipe
activity, L2 hit rate, registers, and merges `dram_bytes_per_launch` (mean of dram__bytes_read.sum + dram__bytes_write.sum over
the captured launches of that kernel) into profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic`."""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--D", type=int, default=4096)
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--summary", default=None)
    ap.add_argument("--labels", default="", help="comma-separated labels of the captured launches, in order")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    labels = [l for l in args.labels.split(",") if l]
    lines, traffic = [], []
    k = 0
    for r in data:
        name = r[col["Kernel Name"]]
        if args.kernel not in name:
            continue
        label = labels[k] if k < len(labels) else "launch %d" % k
        k += 1
        lines.append("## %s   %s" % (label, name[:90]))
        tot = 0.0
        for m in KEEP:
            if m in col and r[col[m]] != "":
                lines.append("   %-88s %s %s" % (m, r[col[m]], units[col[m]]))
                if m.startswith("dram__bytes"):
                    tot += float(r[col[m]].replace(",", "")) * UNIT_BYTES.get(units[col[m]], 1.0)
        traffic.append(tot)
    text = "ncu --set full --clock-control none capture %s (kernel filter %s)\n" % (os.path.basename(args.rep), args.kernel)
    text += "\n".join(lines) + "\n"
    if traffic:
        text += "mean DRAM bytes per launch (read + write): %.1f MB over %d launches\n" % (sum(traffic) / len(traffic) / 1e6, len(traffic))
    print(text)
    if args.summary:
        with open(args.summary, "w") as f:
            f.write(text)
    if traffic:
        p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        tab = {"captures": []}
        if os.path.exists(p):
            with open(p) as f:
                tab = json.load(f)
        tab["captures"] = [c for c in tab["captures"] if not (c["kernel"] == args.kernel and c["D"] == args.D and
                                                                 c["B"] == args.B and c.get("world", 1) == args.world)]
        tab["captures"].append({"kernel": args.kernel, "D": args.D, "B": args.B, "world": args.world,
                                "dram_bytes_per_launch": sum(traffic) / len(traffic), "per_launch": traffic,
                                "source": (args.summary or args.rep) + " (ncu --set full, dram__bytes_read.sum + "
                                          "dram__bytes_write.sum, mean over the captured launches)"})
        with open(p, "w") as f:
            json.dump(tab, f, indent=1)


if __name__ == "__main__":
    main()
