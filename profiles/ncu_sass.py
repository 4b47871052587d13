#!/usr/bin/env python
"""SASS-level view of an ncu report: python profiles/ncu_sass.py report.ncu-rep [min_inst_fraction]
Prints, in program order, every SASS instruction of the first kernel whose execution count is at least the given fraction
(default 0.2) of the most executed instruction, with its sample count and the dominant stall reasons -- the hot loop with
the place where each warp waits.  (`ncu -i ... --page source --csv --print-source sass`.)"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, body, kernels = None, [], 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernels += 1
            if kernels > 1:
                break
            print("kernel:", r[1])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) >= 8:
            body.append(dict(zip(hdr, r)))
    if not body:
        print("no SASS rows")
        return
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    def num(d, k):
        try:
            return int(d.get(k, "0") or 0)
        except ValueError:
            return 0
    mx = max(num(d, "Instructions Executed") for d in body) or 1
    tot_s = sum(num(d, "# Samples") for d in body) or 1
    print(f"instructions {len(body)}, samples {tot_s}, max executions {mx}")
    base = int(body[0]["Address"], 16)
    for d in body:
        ie = num(d, "Instructions Executed")
        if ie < frac * mx:
            continue
        sm = num(d, "# Samples")
        st = sorted(((num(d, c), c[6:]) for c in stall_cols), reverse=True)
        why = " ".join(f"{n}:{c}" for c, n in [(c, n) for n, c in st[:3] if n > 0])
        print(f"{int(d['Address'], 16) - base:5x} {ie / mx:5.2f} {100 * sm / tot_s:5.2f}% {d['Source'].strip()[:70]:70s} {why}")


if __name__ == "__main__":
    main()
