#!/bin/bash
# Round 2, GPU call R: the 17.7 KB decoder (twelve streams per SM): parity subset, then rANS throughput against streams per SM.
tag=${1:-r2r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rans or kat or encode_decode or golden or lossless" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/rans_occupancy.py --frames 8 --chunks 197,394,493,580,592 > gpurun_out/${tag}_occupancy.jsonl 2> gpurun_out/${tag}_occupancy.err
echo "occupancy rc=$?"; cut -c1-330 gpurun_out/${tag}_occupancy.jsonl; tail -3 gpurun_out/${tag}_occupancy.err
