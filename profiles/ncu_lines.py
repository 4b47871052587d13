#!/usr/bin/env python
"""Summarise an ncu report by CUDA source line: python profiles/ncu_lines.py report.ncu-rep [top]
(uses `ncu -i ... --page source --csv --print-source cuda,sass`; needs -lineinfo at compile time)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    fn_count = 0
    for r in rows:
        if r and r[0] == "Function Name":
            fn_count += 1
        if "Instructions Executed" in r:
            hdr = r
            continue
        if fn_count > 1:
            break
        if hdr and len(r) >= len(hdr) - 2 and r[0] not in ("", "Line No"):
            d = dict(zip(hdr, r))
            try:
                lines.append((int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"]), r[0], r[1]))
            except ValueError:
                pass
    tot = sum(x[0] for x in lines) or 1
    tots = sum(x[2] for x in lines) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for ie, te, sm, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"{100 * ie / tot:5.1f}% inst  {100 * sm / tots:5.1f}% samp  thr/inst {te / max(ie, 1):5.1f} | {ln:>4}: {src[:110]}")


if __name__ == "__main__":
    main()
