#!/bin/bash
# One 8-GPU box: the headline bench at 1/2/4/8 GPUs and BASELINE config 4 (10^8 shots of Tanner and bicycle, min-sum layered,
# sharded over the GPUs, one NCCL all-reduce of the counters) at 1/2/4/8 GPUs.  Outputs under gpurun_out/.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1200 -- 'bash benchmarks/scale_all.sh'
O=gpurun_out
mkdir -p $O
: > $O/scale_1_2_4_8.jsonl
: > $O/cfg4.jsonl
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu >> $O/scale_1_2_4_8.jsonl 2>> $O/scale.err
for n in 2 4 8; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 >> $O/scale_1_2_4_8.jsonl 2>> $O/scale.err
done
for code in T bicycle; do
    python benchmarks/cfg4_scaling.py --code $code --shots 100000000 --classes >> $O/cfg4.jsonl 2>> $O/cfg4.err
    for n in 2 4 8; do
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510 + n)) benchmarks/cfg4_scaling.py --code $code --shots 100000000 --classes >> $O/cfg4.jsonl 2>> $O/cfg4.err
    done
done
nvidia-smi topo -m > $O/topo.txt 2>&1
echo done
