#!/bin/bash
# Round 2, GPU call A: measure the switches left open at the end of round 1.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 400 python tools/rans_occupancy.py --frames 8 --chunks 99,197,296,395,592 --envs ALICE_RANS_DEC_SPLIT16=1 \
    > gpurun_out/${tag}_rans_occupancy.jsonl 2> gpurun_out/${tag}_rans_occupancy.err
echo "occupancy rc=$?"; cat gpurun_out/${tag}_rans_occupancy.jsonl; tail -3 gpurun_out/${tag}_rans_occupancy.err
AB_REPS=2 AB_GOLDEN=1 timeout 200 python tools/ab_backend.py base h16 base > gpurun_out/${tag}_ab_backend.jsonl 2> gpurun_out/${tag}_ab_backend.err
echo "ab_backend rc=$?"; cat gpurun_out/${tag}_ab_backend.jsonl
timeout 200 python tools/ab_frontend.py base xysplit base > gpurun_out/${tag}_ab_frontend.jsonl 2> gpurun_out/${tag}_ab_frontend.err
echo "ab_frontend rc=$?"; cat gpurun_out/${tag}_ab_frontend.jsonl
