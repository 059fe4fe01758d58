#!/bin/bash
# One GPU-box call: parity suite, measurements of the 8(f) rows, A/B of the experiment builds.  Outputs -> gpurun_out/
tag=${1:-r30}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 150 python tests/bench_next_rows.py > gpurun_out/${tag}_next_rows.jsonl 2> gpurun_out/${tag}_next_rows.err
cat gpurun_out/${tag}_next_rows.jsonl
timeout 150 python tools/ab_frontend.py base t_rolled t_rolled3 xyu2 xyu4 base > gpurun_out/${tag}_ab_frontend.jsonl 2> gpurun_out/${tag}_ab_frontend.err
cat gpurun_out/${tag}_ab_frontend.jsonl
