#!/usr/bin/env python
"""Turn the outputs of profiles/capture.sh (gpurun_out/) into the tracked summary directory of a round:

    python profiles/summarize.py profiles/r01_v7

  launches.csv            ncu launch list of `bench.py --steps 2 --warmup 3 --no-cpu` (gpu__time_duration.sum per launch)
  launch_shares.txt       time share of every kernel in that list
  ms_kernel_raw.csv       `ncu --page raw --csv` of the full capture of the dominant kernel (2 launches)
  ms_kernel_summary.txt   the metrics the design document quotes (profiles/ncu_summary.py)
  ms_kernel_source_lines.txt  instructions / samples by CUDA source line (profiles/ncu_lines.py)
  bench.json, bench_reference_arm.json, configs.jsonl, pytest_gpu.log
and refresh profiles/traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def main():
    out = os.path.abspath(sys.argv[1])
    os.makedirs(out, exist_ok=True)
    rep = os.path.join(G, "prof_ms_final.ncu-rep")
    for src, dst in (("launches.csv", "launches.csv"), ("bench.log", "bench.json"), ("bench_ref.log", "bench_reference_arm.json"),
                     ("configs.log", "configs.jsonl"), ("pytest_gpu.log", "pytest_gpu.log")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(out, dst))
    # launch shares
    rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows:
        if len(r) == len(hdr) and r[0].isdigit():
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out, "launch_shares.txt"), "w") as f:
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:70]:70s} n={len(v):3d} total={sum(v):10.3f} ms share={sum(v) / tot:.3f}\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(out, "ms_kernel_raw.csv"), "w").write(raw)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(out, "ms_kernel_summary.txt"), "w").write(summ)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_lines.py"), rep, "50"], capture_output=True, text=True).stdout
    open(os.path.join(out, "ms_kernel_source_lines.txt"), "w").write(lines)
    rr = list(csv.reader(raw.splitlines()))
    h, data = rr[0], rr[2:]
    rd = [float(d[h.index("dram__bytes_read.sum")]) for d in data]
    wr = [float(d[h.index("dram__bytes_write.sum")]) for d in data]
    ur, uw = rr[1][h.index("dram__bytes_read.sum")], rr[1][h.index("dram__bytes_write.sum")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = [a * scale[ur] + b * scale[uw] for a, b in zip(rd, wr)]
    name = data[0][h.index("Kernel Name")] if "Kernel Name" in h else "ms_decode_kernel"
    json.dump({"kernel": name, "source": os.path.relpath(os.path.join(out, "ms_kernel_raw.csv"), ROOT) +
               " (ncu --set full, bench.py --steps 2 --warmup 3 --no-cpu, 10^6 shots per launch)",
               "dram_bytes_per_launch": sum(per) / len(per), "per_launch": per},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    try:
        print(open(os.path.join(out, "launch_shares.txt")).read())
        print(summ)
    except BrokenPipeError:
        pass


if __name__ == "__main__":
    main()
