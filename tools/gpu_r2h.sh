#!/bin/bash
# Round 2, GPU call H: two rANS streams per warp: parity on hardware, throughput against streams per SM, default bench.
tag=${1:-r2h}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 1,50,99,197,296,394,592 > gpurun_out/${tag}_rans_occupancy.jsonl 2> gpurun_out/${tag}_rans_occupancy.err
echo "occupancy rc=$?"; cut -c1-330 gpurun_out/${tag}_rans_occupancy.jsonl; tail -3 gpurun_out/${tag}_rans_occupancy.err
timeout 900 python bench.py --steps 4 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -6 gpurun_out/${tag}_bench.err
