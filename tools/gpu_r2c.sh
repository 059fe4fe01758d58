#!/bin/bash
# Round 2, GPU call C: first hardware exposure of the fused front-end kernel: parity, time, ncu.
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "odd_and_edge or fullsize_config1 or fullsize_config2 or shared_workspace or batch_api or baseline_configs_medium" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/frontend_time.py --batch 8 --reps 3 > gpurun_out/${tag}_frontend.jsonl 2> gpurun_out/${tag}_frontend.err
echo "frontend rc=$?"; cat gpurun_out/${tag}_frontend.jsonl; tail -3 gpurun_out/${tag}_frontend.err
for wv in cdf97 cdf53; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fwd_fused -c 1 -o gpurun_out/${tag}_fused_${wv} \
    python tools/prof_chunk.py --frames 64 --chunks 1 --reps 0 --wavelet $wv > gpurun_out/${tag}_ncu_${wv}.log 2>&1
echo "ncu $wv rc=$?"; tail -2 gpurun_out/${tag}_ncu_${wv}.log
done
