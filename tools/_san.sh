cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import oracle_lib as O, gpu_util, parity
g = O.load_golden(sys.argv[1])
got = gpu_util.gpu_render(g["snap"])
print(parity.compare(got, g))
PY
for tool in synccheck racecheck memcheck; do echo "== $tool"; timeout 300 compute-sanitizer --tool $tool python /tmp/one.py cfg0_frame7_320x180 2>&1 | grep -v "^$" | tail -12; done
