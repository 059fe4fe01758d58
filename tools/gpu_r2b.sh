#!/bin/bash
# Round 2, GPU call B: parity of the new decoder layout on hardware + block-shape policy of the rANS launches.
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 197,247,296,345,394,592 --envs ALICE_RANS_FORCE_LPB=1,ALICE_RANS_FORCE_LPB=4 \
    > gpurun_out/${tag}_rans_occupancy.jsonl 2> gpurun_out/${tag}_rans_occupancy.err
echo "occupancy rc=$?"; cat gpurun_out/${tag}_rans_occupancy.jsonl; tail -3 gpurun_out/${tag}_rans_occupancy.err
