#!/bin/bash
# Round 2, GPU call N: bench lines of configs 3 and 5 (one GPU) and the CPU reference arm on full chunks.
tag=${1:-r2n}
mkdir -p gpurun_out
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_cfg2.json 2> gpurun_out/${tag}_bench_reference_cfg2.err
echo "reference rc=$?"; cat gpurun_out/${tag}_bench_reference_cfg2.json | cut -c1-900; tail -2 gpurun_out/${tag}_bench_reference_cfg2.err
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err
echo "cfg3 rc=$?"; cat gpurun_out/${tag}_bench_cfg3.json | cut -c1-1200; tail -3 gpurun_out/${tag}_bench_cfg3.err
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg5_n1.json 2> gpurun_out/${tag}_bench_cfg5_n1.err
echo "cfg5 rc=$?"; cat gpurun_out/${tag}_bench_cfg5_n1.json | cut -c1-1200; tail -3 gpurun_out/${tag}_bench_cfg5_n1.err
