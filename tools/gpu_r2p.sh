#!/bin/bash
# Round 2, GPU call P: the round-end sequence on the final build: -m gpu suite, smoke, default bench (six and nine e2e threads).
tag=${1:-r2p}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 6 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'B',d['config']['chunks_per_gpu_per_step'],'e2e',d['e2e']['value'],d['e2e']['chunks_per_step_per_gpu'],d['e2e']['ms_per_step'],d['e2e']['decoded_matches_oracle_digest'],'bit',d['bit_exact_vs_oracle_digest'])
print(d['stages'])
"; tail -3 gpurun_out/${tag}_bench.err
timeout 900 python bench.py --steps 6 --warmup 3 --e2e-threads 9 --no-cpu-baseline > gpurun_out/${tag}_bench_T9.json 2> gpurun_out/${tag}_bench_T9.err
echo "bench T9 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_T9.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['chunks_per_step_per_gpu'],d['e2e']['ms_per_step'])
"; tail -3 gpurun_out/${tag}_bench_T9.err
