mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/ab_tests.log
for i in 1 2; do timeout 120 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/ab_bench_$i.json 2> gpurun_out/ab_bench.err; python -c "
import json
d=json.load(open('gpurun_out/ab_bench_$i.json')); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],5), 'e2e', round(d['e2e']['value'],1))"; done
timeout 100 python bench.py --no-cpu-baseline --steps 30 --workload cfg4_8k 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('8k ms', round(d['ms_per_step'],4))"
