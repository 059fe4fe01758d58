#!/bin/bash
# Second GPU-box call of the round: parity suite at the final defaults, 8(f) rows, the full default bench, then (time
# permitting) a launch list of a 4-chunk bench step and one full ncu capture of the rolled t kernel.
tag=${1:-r31}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 150 python tests/bench_next_rows.py > gpurun_out/${tag}_next_rows.jsonl 2> gpurun_out/${tag}_next_rows.err
cat gpurun_out/${tag}_next_rows.jsonl
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_fwd_t_quant -c 1 -o gpurun_out/${tag}_tquant \
    python tools/prof_chunk.py --frames 64 --chunks 1 --reps 0 > gpurun_out/${tag}_ncu_tquant.log 2>&1
echo "ncu tquant rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --chunks 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
