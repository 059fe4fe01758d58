#!/bin/bash
# Round 2, GPU call J: rANS decoder chain variants (A/B/C), each through the occupancy sweep at 1 / 4 / 8 streams per SM,
# with a parity spot check per build.
tag=${1:-r2j}
mkdir -p gpurun_out
for v in base spec0 spec2; do
  lib=alice-codec_b200/lib/libalice_codec.so
  [ $v != base ] && lib=alice-codec_b200/lib/libalice_codec_${v}.so
  timeout 200 python tests/check_variant_gpu.py $lib > gpurun_out/${tag}_check_${v}.json 2>&1
  echo "check $v rc=$?"; tail -1 gpurun_out/${tag}_check_${v}.json | cut -c1-300
  timeout 200 python tools/rans_occupancy.py --frames 8 --chunks 1,197,394 --lib $lib > gpurun_out/${tag}_occ_${v}.jsonl 2> gpurun_out/${tag}_occ_${v}.err
  echo "occ $v rc=$?"; cut -c1-330 gpurun_out/${tag}_occ_${v}.jsonl; tail -2 gpurun_out/${tag}_occ_${v}.err
done
