#!/bin/bash
# Round 2, GPU call D: fused front-end after the copy-issue fix (A/B against the two-kernel path), new bench.py.
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 300 python tools/frontend_time.py --batch 8 --reps 2 --decode > gpurun_out/${tag}_frontend_fused.jsonl 2> gpurun_out/${tag}_frontend_fused.err
echo "fused rc=$?"; cat gpurun_out/${tag}_frontend_fused.jsonl; tail -3 gpurun_out/${tag}_frontend_fused.err
ALICE_FWD_FUSED=0 timeout 300 python tools/frontend_time.py --batch 8 --reps 2 > gpurun_out/${tag}_frontend_twokernel.jsonl 2> gpurun_out/${tag}_frontend_twokernel.err
echo "two-kernel rc=$?"; cat gpurun_out/${tag}_frontend_twokernel.jsonl; tail -3 gpurun_out/${tag}_frontend_twokernel.err
timeout 300 python bench.py --chunks 12 --steps 1 --warmup 3 --no-cpu-baseline --verbose > gpurun_out/${tag}_bench_small.json 2> gpurun_out/${tag}_bench_small.err
echo "bench small rc=$?"; cat gpurun_out/${tag}_bench_small.json; tail -5 gpurun_out/${tag}_bench_small.err
timeout 600 python bench.py --steps 2 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -12 gpurun_out/${tag}_bench.err
