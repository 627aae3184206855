cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import oracle_lib as O, gpu_util, parity
g = O.load_golden(sys.argv[1])
got = gpu_util.gpu_render(g["snap"], stats=False)
print("class mismatches", int((got["cls"] != g["cls"]).sum()), "step mismatches", int((got["steps"] != g["steps"]).sum()), got["steps"][119,103], g["steps"][119,103])
PY
export BH8_LIB_PATH=build/variants/dbg.so
python /tmp/one.py cfg0_frame7_320x180
