#!/bin/bash
# Round 2, GPU call T: encoder with the byte decision first (throughput against streams per SM), then one ncu --set full
# capture of the rANS kernels at eight (26.4 KB decoder) and twelve (17.7 KB decoder) streams per SM.
tag=${1:-r2t}
mkdir -p gpurun_out
v=old8
timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 197,394 --lib alice-codec_b200/lib/libalice_codec_$v.so > gpurun_out/${tag}_occ_$v.jsonl 2> gpurun_out/${tag}_occ_$v.err
echo "$v rc=$?"; python - <<PY
import json
for l in open("gpurun_out/${tag}_occ_$v.jsonl"):
    d=json.loads(l); print(" ", d["streams_per_sm"], "enc", d["enc_msym_s_per_lane_if_all_resident"], d["enc_msym_s_per_sm"], "dec", d["dec_msym_s_per_lane_if_all_resident"], d["dec_msym_s_per_sm"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rans -c 2 -o gpurun_out/${tag}_rans_old8_8persm -f \
  python tools/rans_occupancy.py --frames 4 --chunks 394 --lib alice-codec_b200/lib/libalice_codec_old8.so > gpurun_out/${tag}_ncu_old8.log 2>&1
echo "ncu old8 rc=$?"; tail -2 gpurun_out/${tag}_ncu_old8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rans -c 2 -o gpurun_out/${tag}_rans_new12_12persm -f \
  python tools/rans_occupancy.py --frames 4 --chunks 592 --lib alice-codec_b200/lib/libalice_codec_new12.so > gpurun_out/${tag}_ncu_new12.log 2>&1
echo "ncu new12 rc=$?"; tail -2 gpurun_out/${tag}_ncu_new12.log
ls -la gpurun_out/*.ncu-rep
