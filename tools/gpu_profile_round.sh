#!/bin/bash
# One GPU call that produces the evidence set of a round under gpurun_out/<tag>_*:
#   full ncu capture of the render kernel (read here with tools/ncu_summary.py / ncu_function_table.py /
#   instr_model.py), the launch list of a short bench run, the GPU test log, the bench line.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/${tag}_gpu_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bh8_render_kernel -s 30 -c 1 -o gpurun_out/${tag}_prof -f \
  python bench.py --kernel-only --steps 20 --warmup 5 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
  python bench.py --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/${tag}_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; echo "ref rc=$?"
