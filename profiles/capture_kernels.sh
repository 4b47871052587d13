#!/bin/bash
# Per-kernel ncu captures (one GPU): every kernel of the library on the configuration that exercises it, each only after the
# same command has exited 0 without ncu.  The reports are summarised ON THE BOX (profiles/summarize_kernels.py -> text under
# gpurun_out/ksum/, copy it to profiles/r02_xx/) and then deleted: a dozen full reports exceed what gpurun copies back.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/capture_kernels.sh [tag ...]'
set -u
O=gpurun_out
mkdir -p $O
cap() {   # tag, kernel regex, launches to skip, run_one arguments...
    local tag=$1 rx=$2 skip=$3; shift 3
    if [ -n "${ONLY:-}" ] && [[ " $ONLY " != *" $tag "* ]]; then return; fi
    python benchmarks/run_one.py "$@" > $O/k_$tag.plain.json 2> $O/k_$tag.err || { echo "$tag: plain run failed" >> $O/capture_kernels.log; return; }
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 2 -f -o $O/k_$tag python benchmarks/run_one.py "$@" > $O/k_$tag.ncu.log 2>&1
    echo "$tag: rc=$?" >> $O/capture_kernels.log
}
ONLY="$*"
: > $O/capture_kernels.log
cap ms_headline   ms_decode  2 --code LP118_0 --dec MS --sched L --p 0.05 --shots 1000000
cap ms_team       ms_decode  2 --code LP118_2 --dec MS --sched L --p 0.05 --shots 200000
cap ms_serial     ms_decode  2 --code LP118_2 --dec MS --sched S --p 0.05 --shots 200000
cap ms_bicycle    ms_sub     2 --code bicycle --dec MS --sched L --p 0.03 --shots 200000
cap ms_lp04       ms_decode  2 --code LP04_0 --dec MS --sched L --p 0.05 --shots 1000000
cap ms_flooding   ms_decode  2 --code LP118_0 --dec MS --sched F --p 0.05 --shots 200000
cap bp            bp_decode  2 --code LP118_0 --dec BP --sched F --p 0.05 --iters 100 --shots 50000
cap osd1          osd_       2 --code LP118_0 --dec MS --sched L --p 0.10 --osd 0 --shots 50000
cap osd2          osd_       2 --code LP118_2 --dec MS --sched S --p 0.05 --osd 10 --shots 100000
cap bf            bf_sparse  2 --code LP118_0 --dec BF --p 0.02 --shots 100000
cap ng            ng_decode  2 --code LP118_0 --dec NG --p 0.02 --shots 100000
cap classify      classify   1 --code LP118_0 --dec MS --sched L --p 0.05 --shots 1000000
cap sample        sample_kernel 0 --code LP118_0 --dec MS --sched L --p 0.05 --shots 1000000
python profiles/summarize_kernels.py $O/ksum > $O/summarize_kernels.log 2>&1
for t in ms_headline ms_team ms_serial ms_bicycle ms_lp04 bp osd2; do
    [ -f $O/k_$t.ncu-rep ] && python profiles/ncu_lines.py $O/k_$t.ncu-rep 40 > $O/ksum/${t}_source_lines.txt 2>/dev/null
    [ -f $O/k_$t.ncu-rep ] && python profiles/ncu_sass.py $O/k_$t.ncu-rep 0.2 > $O/ksum/${t}_sass.txt 2>/dev/null
done
[ -f $O/k_ms_headline.ncu-rep ] && ncu -i $O/k_ms_headline.ncu-rep --page raw --csv > $O/ksum/ms_headline_raw.csv 2>/dev/null
rm -f $O/k_*.ncu-rep
echo done >> $O/capture_kernels.log
