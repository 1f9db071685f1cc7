#!/bin/bash
# N-GPU A/B of the two gradient all-reduce forms (run under `gpurun --gpus N`): tools/bench_8gpu_allreduce.sh N
N=${1:-8}
run() { name=$1; shift; PDDM_BENCH_WATCHDOG=100 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --no-cpu-baseline --no-sampling "$@" > gpurun_out/s8_${N}gpu_$name.json 2> gpurun_out/s8_${N}gpu_$name.err; echo "$name rc=$? $(cut -c1-200 gpurun_out/s8_${N}gpu_$name.json | tail -1)"; }
run overlap --allreduce overlap
run flat --allreduce flat
