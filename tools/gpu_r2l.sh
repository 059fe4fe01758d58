#!/bin/bash
# Round 2, GPU call L: fused front-end after the t-phase trims; e2e host-thread count.
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 300 python tools/frontend_time.py --batch 8 --reps 2 --decode > gpurun_out/${tag}_fused.jsonl 2> gpurun_out/${tag}_fused.err
echo "fused rc=$?"; cut -c1-330 gpurun_out/${tag}_fused.jsonl; tail -3 gpurun_out/${tag}_fused.err
for T in 9 12; do
timeout 900 python bench.py --steps 6 --warmup 3 --e2e-threads $T --no-cpu-baseline > gpurun_out/${tag}_bench_T$T.json 2> gpurun_out/${tag}_bench_T$T.err
echo "bench T=$T rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_T$T.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['chunks_per_step_per_gpu'],d['e2e']['ms_per_step'],d['e2e']['decoded_matches_oracle_digest'])
"; tail -3 gpurun_out/${tag}_bench_T$T.err
done
