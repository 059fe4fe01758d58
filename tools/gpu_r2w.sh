#!/bin/bash
# Round 2, GPU call W: final build (decoder keeps only the shifted state, prefetched window refills; encoder entries carry
# their shift): rANS lanes against streams per SM, the whole -m gpu suite, smoke, the default bench.
tag=${1:-r2w}
mkdir -p gpurun_out
timeout 300 python tools/rans_occupancy.py --frames 8 --chunks 197,394 > gpurun_out/${tag}_occupancy.jsonl 2> gpurun_out/${tag}_occupancy.err
echo "occupancy rc=$?"; python - <<PY
import json
for l in open("gpurun_out/${tag}_occupancy.jsonl"):
    d=json.loads(l); print(" ", d["streams_per_sm"], "enc", d["enc_msym_s_per_lane_if_all_resident"], d["enc_msym_s_per_sm"], "dec", d["dec_msym_s_per_lane_if_all_resident"], d["dec_msym_s_per_sm"])
PY
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 6 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'B',d['config']['chunks_per_gpu_per_step'],'ms',d['ms_per_step'],'bit',d['bit_exact_vs_oracle_digest'])
print(d['stages'])
e=d['e2e']; print('e2e',e['value'],e['chunks_per_step_per_gpu'],e['ms_per_step'],e['decoded_matches_oracle_digest'])
print(d['cpu_baseline'])
"; tail -4 gpurun_out/${tag}_bench.err
