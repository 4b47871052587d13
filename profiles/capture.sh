#!/bin/bash
# One gpurun call: GPU tests, smoke, headline bench (own arm + reference arm), configuration sweep, then the ncu launch list of the
# SAME bench command (gpu__time_duration per launch).  Per-kernel full captures: profiles/capture_kernels.sh.  Outputs: gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash profiles/capture.sh'
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python bench.py > $O/bench.log 2> $O/bench.err
python bench.py --impl reference > $O/bench_ref.log 2>> $O/bench.err
python benchmarks/configs.py > $O/configs.log 2> $O/configs.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
echo done
