#!/bin/bash
# What is one more resident shot per SM worth?  Headline configuration with the shots per CTA capped at 16 .. 20 (QLDPC_SHOTS_CAP).
# Recorded in DESIGN.md section 9 (the quasi-cyclic addressing decision rests on it): gpurun -- 'bash benchmarks/occupancy_sweep.sh'
for cap in 16 17 18 19 20; do
    QLDPC_SHOTS_CAP=$cap python benchmarks/run_one.py --code LP118_0 --dec MS --sched L --p 0.05 --shots 1000000 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('shots_per_cta', d['info']['shots_per_cta'], 'shots_per_s', round(d['shots_per_s']))"
done
