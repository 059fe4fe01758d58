#!/bin/bash
# Round 2, GPU call E: whole -m gpu suite after the engine changes, default bench with the tight payload arena and six
# e2e host threads.
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --verbose > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -30 gpurun_out/${tag}_bench.err
