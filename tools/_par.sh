for lib in default build/variants/flat.so build/variants/constregs.so; do
  if [ "$lib" = default ]; then unset BH8_LIB_PATH; else export BH8_LIB_PATH=$lib; fi
  echo "== $lib"; timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
done
