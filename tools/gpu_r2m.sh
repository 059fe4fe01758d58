#!/bin/bash
# Round 2, GPU call M: encoder with multiply-high instead of a register-count shift (parity spot check + throughput sweep).
tag=${1:-r2m}
mkdir -p gpurun_out
timeout 200 python tests/check_variant_gpu.py alice-codec_b200/lib/libalice_codec.so > gpurun_out/${tag}_check.json 2>&1
echo "check rc=$?"; tail -1 gpurun_out/${tag}_check.json | cut -c1-300
timeout 200 python tools/rans_occupancy.py --frames 8 --chunks 1,197,394 > gpurun_out/${tag}_occ.jsonl 2> gpurun_out/${tag}_occ.err
echo "occ rc=$?"; cut -c1-330 gpurun_out/${tag}_occ.jsonl; tail -2 gpurun_out/${tag}_occ.err
