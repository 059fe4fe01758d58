#!/bin/bash
# Round 2, GPU call AB: the e2e leg with four host threads, second sample.
tag=${1:-r2ab}
mkdir -p gpurun_out
timeout 420 python bench.py --steps 1 --warmup 3 --e2e-steps 6 --e2e-threads 4 --no-cpu-baseline > gpurun_out/${tag}_bench_T4.json 2> gpurun_out/${tag}_bench_T4.err
echo "T=4 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_T4.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value',d['value'],'e2e',e['value'],e['chunks_per_step_per_gpu'],e['ms_per_step'],e['decoded_matches_oracle_digest'])
"; tail -2 gpurun_out/${tag}_bench_T4.err
