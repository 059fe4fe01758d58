#!/bin/bash
# Round 2, GPU call X: ncu warp-state / scheduler / pipe sections of k_rans_decode (the dominant kernel) at eight streams per SM
# on the final build, then smoke and a short device-resident bench of the committed tree.
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 420 ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --clock-control none -k regex:k_rans_decode -c 1 -o gpurun_out/${tag}_rans_decode_8persm -f \
  python tools/rans_occupancy.py --frames 2 --chunks 394 > gpurun_out/${tag}_ncu_dec.log 2>&1
echo "ncu dec rc=$?"; tail -2 gpurun_out/${tag}_ncu_dec.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'B',d['config']['chunks_per_gpu_per_step'],'ms',d['ms_per_step'],'bit',d['bit_exact_vs_oracle_digest'])
"; tail -2 gpurun_out/${tag}_bench.err
