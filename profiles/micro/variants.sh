#!/bin/bash
# A/B of build variants: bash profiles/micro/variants.sh TAG lib1.so lib2.so ...   (bench.py --no-cpu --no-strong per library)
TAG=$1; shift
for lib in default "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" = default ]; then unset RN_LIB; else export RN_LIB=$PWD/$lib; fi
  python bench.py --no-cpu --no-strong --steps 100 > gpurun_out/${TAG}_$name.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_$name.json"))
print("%-14s value %.1f ms %.4f stages %s" % ("$name", d["value"], d["ms_per_step"], [round(x * 1e3, 1) for x in d["stage_ms"]["launch_ms"]]))
PY
done
