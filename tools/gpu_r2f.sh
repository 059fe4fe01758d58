#!/bin/bash
# Round 2, GPU call F: fused back-end: parity on hardware, A/B against the two-kernel back-end, ncu, default bench.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/frontend_time.py --batch 8 --reps 2 --decode > gpurun_out/${tag}_fused.jsonl 2> gpurun_out/${tag}_fused.err
echo "fused rc=$?"; cat gpurun_out/${tag}_fused.jsonl; tail -3 gpurun_out/${tag}_fused.err
ALICE_INV_FUSED=0 ALICE_FWD_FUSED=0 timeout 300 python tools/frontend_time.py --batch 8 --reps 2 --decode > gpurun_out/${tag}_twokernel.jsonl 2> gpurun_out/${tag}_twokernel.err
echo "two-kernel rc=$?"; cat gpurun_out/${tag}_twokernel.jsonl; tail -3 gpurun_out/${tag}_twokernel.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_inv_fused -c 1 -o gpurun_out/${tag}_inv_fused_cdf97 \
    python tools/prof_chunk.py --frames 64 --chunks 1 --reps 0 --wavelet cdf97 > gpurun_out/${tag}_ncu_inv.log 2>&1
echo "ncu inv rc=$?"; tail -2 gpurun_out/${tag}_ncu_inv.log
timeout 900 python bench.py --steps 2 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -8 gpurun_out/${tag}_bench.err
