for ri in 2 1; do
export BH8_SINK_RESTART_INTERVAL=$ri
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpeg_ -c 8 --csv --log-file gpurun_out/jpeg_t.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/jpeg_t.csv")) if len(r)>10]
h=rows[0]
print("ri", $ri)
for r in rows[3:6]:
    d=dict(zip(h,r)); print(" ", d["Kernel Name"][:30], d["Metric Value"], d["Metric Unit"])
PY
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['sink']; print('  fps sync %.0f pipelined %.0f encode_ms %.4f bytes %.0f' % (s['frames_per_s'], s['pipelined_bh8_sink_submit']['frames_per_s'], s['encode_device_ms_per_frame'], s['jpeg_bytes_per_frame']))"
done
