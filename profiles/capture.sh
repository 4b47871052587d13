#!/bin/bash
# One gpurun call: GPU tests, headline bench (own arm + reference arm), configuration sweep, then the two ncu passes of the
# SAME bench command (launch list with gpu__time_duration, full capture of the dominant kernel).  Outputs: gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash profiles/capture.sh'
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py > $O/bench.log 2> $O/bench.err
python bench.py --impl reference > $O/bench_ref.log 2>> $O/bench.err
python benchmarks/configs.py > $O/configs.log 2> $O/configs.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ms_decode -s 6 -c 2 -f -o $O/prof_ms_final $CMD > $O/ncu2.log 2>&1
echo done
