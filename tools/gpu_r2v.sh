#!/bin/bash
# Round 2, GPU call V: twelve streams per SM end to end: parity subset, then the default bench with the streaming memory plan.
tag=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "stream_device or submit_collect or shifted or batch or rans or kat or golden or arena" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 4 --no-cpu-baseline --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'B',d['config']['chunks_per_gpu_per_step'],'ms',d['ms_per_step'],'bit',d['bit_exact_vs_oracle_digest'])
print(d['stages'])
e=d['e2e']; print('e2e',e['value'],e['chunks_per_step_per_gpu'],e['ms_per_step'],e['decoded_matches_oracle_digest'])
"; tail -8 gpurun_out/${tag}_bench.err
