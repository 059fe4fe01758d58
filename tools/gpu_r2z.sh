#!/bin/bash
# Round 2, GPU call Z: decoder with a table-free step for symbol 0 (slots below freq[0]: freq and slot - cum are known without
# the table read): rANS throughput against streams per SM; then the e2e leg with 4 and 5 host threads.
tag=${1:-r2z}
mkdir -p gpurun_out
for v in top; do
  timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 197,394 --lib alice-codec_b200/lib/libalice_codec_$v.so > gpurun_out/${tag}_occ_$v.jsonl 2> gpurun_out/${tag}_occ_$v.err
  echo "$v rc=$?"; python - <<PY
import json
for l in open("gpurun_out/${tag}_occ_$v.jsonl"):
    d=json.loads(l); print(" ", d["streams_per_sm"], "enc", d["enc_msym_s_per_lane_if_all_resident"], d["enc_msym_s_per_sm"], "dec", d["dec_msym_s_per_lane_if_all_resident"], d["dec_msym_s_per_sm"])
PY
  tail -2 gpurun_out/${tag}_occ_$v.err
done
timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 394 --wavelet cdf53 --quality 90 --lib alice-codec_b200/lib/libalice_codec_top.so 2>&1 | cut -c1-330
timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 394 --wavelet cdf53 --quality 90 2>&1 | cut -c1-330
