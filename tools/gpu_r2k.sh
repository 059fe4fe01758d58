#!/bin/bash
# Round 2, GPU call K: default bench with the e2e batches on the small-shared-memory kernels, then with the fused ones.
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 6 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json | cut -c1-1500; tail -4 gpurun_out/${tag}_bench.err
timeout 900 python bench.py --steps 6 --warmup 3 --e2e-fused --no-cpu-baseline --verbose > gpurun_out/${tag}_bench_e2efused.json 2> gpurun_out/${tag}_bench_e2efused.err
echo "bench e2e-fused rc=$?"; cat gpurun_out/${tag}_bench_e2efused.json | cut -c1-1500; tail -4 gpurun_out/${tag}_bench_e2efused.err
