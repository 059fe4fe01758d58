#!/bin/bash
# Round 2, GPU call U: fewer shared-memory wavefronts per symbol in the serial rANS loops (decoder: 8-byte look-ahead topped up
# every other symbol, slots stored four at a time; encoder: one 16-byte entry per symbol instead of entry + shift word, states
# stored four at a time): throughput against streams per SM.
tag=${1:-r2u}
mkdir -p gpurun_out
for v in d1 d1e1 d1e2 n12d1e2; do
  c=197,394; [ "$v" = n12d1e2 ] && c=394,592
  timeout 300 python tools/rans_occupancy.py --frames 8 --chunks $c --lib alice-codec_b200/lib/libalice_codec_$v.so > gpurun_out/${tag}_occ_$v.jsonl 2> gpurun_out/${tag}_occ_$v.err
  echo "$v rc=$?"; python - <<PY
import json
for l in open("gpurun_out/${tag}_occ_$v.jsonl"):
    d=json.loads(l); print(" ", d["streams_per_sm"], "enc", d["enc_msym_s_per_lane_if_all_resident"], d["enc_msym_s_per_sm"], "dec", d["dec_msym_s_per_lane_if_all_resident"], d["dec_msym_s_per_sm"])
PY
  tail -2 gpurun_out/${tag}_occ_$v.err
done
