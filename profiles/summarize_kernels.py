#!/usr/bin/env python
"""Tracked text summaries of the per-kernel captures of profiles/capture_kernels.sh:

    python profiles/summarize_kernels.py profiles/r02_v2

For every gpurun_out/k_<tag>.ncu-rep: <tag>_summary.txt (kernel name, launch geometry, duration, instructions, issue-active,
pipe utilisation, shared-memory wavefronts / bank conflicts, DRAM bytes, stall reasons -- the format of ms_kernel_summary.txt)
and <tag>_plain.json (the same command's result without ncu: shots/s, iterations)."""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
EXTRA = ["launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
         "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
         "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg"]


def main():
    out = os.path.abspath(sys.argv[1])
    os.makedirs(out, exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(G, "k_*.ncu-rep"))):
        tag = os.path.basename(rep)[2:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            print(tag, "empty report")
            continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        with open(os.path.join(out, f"{tag}_summary.txt"), "w") as f:
            if "Kernel Name" in hdr:
                f.write("kernel: " + " | ".join(sorted(set(d[hdr.index("Kernel Name")] for d in data))) + "\n")
            cmd = open(os.path.join(G, f"k_{tag}.plain.json")).read().strip() if os.path.exists(os.path.join(G, f"k_{tag}.plain.json")) else ""
            f.write("same command without ncu: " + cmd + "\n")
            f.write(summ)
            for k in EXTRA:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:90s} {units[i]:14s} {[d[i] for d in data]}\n")
        if os.path.exists(os.path.join(G, f"k_{tag}.plain.json")):
            shutil.copy(os.path.join(G, f"k_{tag}.plain.json"), os.path.join(out, f"{tag}_plain.json"))
        if tag == "ms_headline":
            # what bench.py quotes beside the (notional) HBM model of the dominant kernel
            def col(k):
                return [float(d[hdr.index(k)].replace(",", "")) for d in data] if k in hdr else []
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            per = [a * scale[ur] + b * scale[uw] for a, b in zip(rd, wr)]
            name = data[0][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "ms_decode_kernel"
            src = f"profiles/<round dir>/{tag}_summary.txt (ncu --set full, benchmarks/run_one.py LP118_0 MS-L p=0.05, 10^6 shots per launch)"
            json.dump({"kernel": name, "source": src, "dram_bytes_per_launch": sum(per) / len(per), "per_launch": per},
                      open(os.path.join(out, "traffic.json"), "w"), indent=1)
            ia = col("smsp__issue_active.avg.pct_of_peak_sustained_active")
            sw = col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
            wi = col("smsp__inst_executed.sum")
            json.dump({"kernel": name, "source": src, "issue_active_pct": sum(ia) / len(ia), "smem_wavefronts_pct_of_peak": sum(sw) / len(sw),
                       "warp_instructions_per_launch": sum(wi) / len(wi),
                       "bound": "instruction issue and shared-memory wavefronts (state is on chip); HBM traffic is ~1e-3 of the algorithmic bytes"},
                      open(os.path.join(out, "on_chip.json"), "w"), indent=1)
        print(tag, "ok")


if __name__ == "__main__":
    main()
