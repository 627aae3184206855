#!/bin/bash
# Kernel-only bench of every tuning variant under build/variants/ (built by blackhole_8_b200.build with -D
# defines) plus run-time knobs, one JSON line each.  usage (GPU box): bash tools/gpu_sweep.sh [workload]
wl=${1:-cfg1_spin}
run() { "$@" python bench.py --kernel-only --steps 200 --warmup 10 --workload $wl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-60s %9.1f Mrays/s  %.5f ms' % ('$*', d['value'], d['ms_per_step']))"; }
run env
run env
for so in build/variants/*.so; do run env BH8_LIB_PATH=$so; done
run env BH8_RESOLVE_WAIT=1
run env BH8_RESOLVE_WAIT=3
