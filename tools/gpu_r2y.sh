#!/bin/bash
# Round 2, GPU call Y: e2e leg against the number of host threads (batches in flight) on the final build.
tag=${1:-r2y}
mkdir -p gpurun_out
for T in 4 5; do
  timeout 600 python bench.py --steps 1 --warmup 3 --e2e-steps 6 --e2e-threads $T --no-cpu-baseline > gpurun_out/${tag}_bench_T$T.json 2> gpurun_out/${tag}_bench_T$T.err
  echo "T=$T rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_T$T.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value',d['value'],'e2e',e['value'],e['chunks_per_step_per_gpu'],e['ms_per_step'],e['decoded_matches_oracle_digest'])
"; tail -2 gpurun_out/${tag}_bench_T$T.err
done
