#!/bin/bash
# One GPU call per kernel experiment: parity of the small fixtures, kernel-only rate (cfg 1 and 8K / nstep 200),
# and the few ncu counters that tell where the time goes (executed warp instructions, issue slots, FP64 pipe).
# usage (GPU box): bash tools/gpu_kernel_iter.sh TAG [full]     -- libs: default + build/variants/*.so
tag=${1:-iter}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_random_scenes.py -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1
echo "parity rc=$? $(tail -1 gpurun_out/${tag}_parity.log)"
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__warps_eligible.avg.per_cycle_active
libs="default"
for so in build/variants/*.so; do [ -e "$so" ] && libs="$libs $so"; done
for lib in $libs; do
  if [ "$lib" = default ]; then unset BH8_LIB_PATH; else export BH8_LIB_PATH=$lib; fi
  name=$(basename $lib .so)
  for rep in 1 2; do
    python bench.py --kernel-only --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-40s cfg1 %9.1f Mrays/s  %.5f ms' % ('$name', d['value'], d['ms_per_step']))"
  done
  python bench.py --kernel-only --steps 5 --warmup 3 --workload cfg4_8k 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-40s 8k   %9.1f Mrays/s  %.5f ms' % ('$name', d['value'], d['ms_per_step']))"
  timeout 200 ncu --metrics $M --clock-control none -k regex:bh8_render_kernel -s 20 -c 1 --csv --log-file gpurun_out/${tag}_${name}_ncu.csv \
    python bench.py --kernel-only --steps 20 --warmup 5 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_${name}_ncu.csv")) if len(r)>10]
hdr=rows[0]; 
for r in rows[1:]:
    d=dict(zip(hdr,r)); print("  ncu %-60s %s %s" % (d.get("Metric Name"), d.get("Metric Value"), d.get("Metric Unit")))
PY
done
if [ "$2" = full ]; then
  unset BH8_LIB_PATH
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bh8_render_kernel -s 30 -c 1 -o gpurun_out/${tag}_prof -f \
    python bench.py --kernel-only --steps 20 --warmup 5 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu full rc=$?"
fi
