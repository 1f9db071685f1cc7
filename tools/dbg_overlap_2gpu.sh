#!/bin/bash
# 2-GPU check of the overlapped arena all-reduce (run under `gpurun --gpus 2`): DP tests, then bench flat vs overlap
timeout 330 python -m pytest tests/test_dp_gpu.py -x -q -m gpu -s > gpurun_out/s7_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_dp.log
grep -i "dp overlap\|parity\|passed\|failed\|rc=\|Error\|File" gpurun_out/s7_dp.log | head -40
run() { name=$1; shift; PDDM_BENCH_WATCHDOG=90 timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --no-cpu-baseline --no-sampling "$@" > gpurun_out/s7_$name.json 2> gpurun_out/s7_$name.err; echo "$name rc=$? $(cut -c1-200 gpurun_out/s7_$name.json | tail -1)"; }
run flat --allreduce flat
run own8 --allreduce overlap --sm-reserve 8
run own4 --allreduce overlap --sm-reserve 4
run flat_b --allreduce flat
run own8_b --allreduce overlap --sm-reserve 8
