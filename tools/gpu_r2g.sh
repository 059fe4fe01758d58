#!/bin/bash
# Round 2, GPU call G: whole suite with the new tests, fused back-end after the load fix, cfg4 / cfg1 bench lines,
# single-chunk latency through the reference ABI, standalone Wavelet3D timing.
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/${tag}_pytest.log
timeout 300 python tools/frontend_time.py --batch 8 --reps 2 --decode > gpurun_out/${tag}_fused.jsonl 2> gpurun_out/${tag}_fused.err
echo "fused rc=$?"; cat gpurun_out/${tag}_fused.jsonl | cut -c1-400; tail -3 gpurun_out/${tag}_fused.err
timeout 200 python tools/wavelet_time.py > gpurun_out/${tag}_wavelet3d.jsonl 2> gpurun_out/${tag}_wavelet3d.err
echo "wavelet rc=$?"; cat gpurun_out/${tag}_wavelet3d.jsonl; tail -3 gpurun_out/${tag}_wavelet3d.err
timeout 300 python tools/latency_abi.py > gpurun_out/${tag}_latency.json 2> gpurun_out/${tag}_latency.err
echo "latency rc=$?"; cat gpurun_out/${tag}_latency.json; tail -3 gpurun_out/${tag}_latency.err
timeout 400 python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err
echo "cfg4 rc=$?"; cat gpurun_out/${tag}_bench_cfg4.json; tail -3 gpurun_out/${tag}_bench_cfg4.err
timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg1.json 2> gpurun_out/${tag}_bench_cfg1.err
echo "cfg1 rc=$?"; cat gpurun_out/${tag}_bench_cfg1.json; tail -3 gpurun_out/${tag}_bench_cfg1.err
