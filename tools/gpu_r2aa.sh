#!/bin/bash
# Round 2, GPU call AA: one scratch volume per device (leased per chunk) instead of one per engine: parity of the two-kernel paths
# and the batch calls, then the default bench (six e2e threads now share one scratch volume: more chunks in flight).
tag=${1:-r2aa}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "encode_decode or batch or shifted or submit or stream or arena or foreign or errors or kat" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --e2e-steps 6 --no-cpu-baseline --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'B',d['config']['chunks_per_gpu_per_step'],'ms',d['ms_per_step'],'bit',d['bit_exact_vs_oracle_digest'])
e=d['e2e']; print('e2e',e['value'],e['chunks_per_step_per_gpu'],e['ms_per_step'],e['decoded_matches_oracle_digest'])
"; tail -6 gpurun_out/${tag}_bench.err
