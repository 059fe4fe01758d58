#!/bin/bash
# Round 2, GPU call S: rANS instruction diet (decoder keeps only the shifted state; encoder decides the byte first), in the
# 26.4 KB (eight per SM) and the 17.7 KB (twelve per SM) decoder layouts: throughput against streams per SM.
tag=${1:-r2s}
mkdir -p gpurun_out
for v in old8 old8m old8e0 new12 new12m; do
  c=197,394; [ "${v:0:5}" = new12 ] && c=197,394,592
  timeout 300 python tools/rans_occupancy.py --frames 8 --chunks $c --lib alice-codec_b200/lib/libalice_codec_$v.so > gpurun_out/${tag}_occ_$v.jsonl 2> gpurun_out/${tag}_occ_$v.err
  echo "$v rc=$?"; python - <<PY
import json
for l in open("gpurun_out/${tag}_occ_$v.jsonl"):
    d=json.loads(l); print(" ", d["streams_per_sm"], "enc", d["enc_msym_s_per_lane_if_all_resident"], d["enc_msym_s_per_sm"], "dec", d["dec_msym_s_per_lane_if_all_resident"], d["dec_msym_s_per_sm"])
PY
  tail -2 gpurun_out/${tag}_occ_$v.err
done
