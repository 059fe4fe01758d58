#!/bin/bash
# Round 2, GPU call I: ncu captures of the final kernels (one chunk each) + the launch list of a 4-chunk bench step.
tag=${1:-r2i}
mkdir -p gpurun_out
for wv in cdf97 cdf53; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_fused|k_inv_fused" -c 2 -o gpurun_out/${tag}_wave_${wv} \
    python tools/prof_chunk.py --frames 64 --chunks 1 --reps 0 --wavelet $wv --quality $([ $wv = cdf97 ] && echo 80 || echo 90) > gpurun_out/${tag}_ncu_wave_${wv}.log 2>&1
echo "ncu wave $wv rc=$?"; tail -2 gpurun_out/${tag}_ncu_wave_${wv}.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_rans_encode|k_rans_decode" -c 2 -o gpurun_out/${tag}_rans \
    python tools/prof_chunk.py --frames 4 --chunks 1 --reps 0 > gpurun_out/${tag}_ncu_rans.log 2>&1
echo "ncu rans rc=$?"; tail -2 gpurun_out/${tag}_ncu_rans.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --chunks 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/${tag}_ncu_bench.log
