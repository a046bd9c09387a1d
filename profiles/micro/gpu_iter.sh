#!/bin/bash
# One GPU iteration: parity tests, the default bench line, and (optionally) one ncu --set full capture of the output kernel.
#   bash profiles/micro/gpu_iter.sh TAG [ncu]
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "stages", [round(x * 1e3, 1) for x in d["stage_ms"]["launch_ms"]], "frac", d["roofline"]["frac"], d["roofline"]["frac_step"])
print("strong", {k: round(v["value"], 1) for k, v in d["strong"].items()})
PY
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_march -s 1 -c 1 -f -o gpurun_out/${TAG}_march python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log
fi
if [ "$2" = "vol" ] || [ "$3" = "vol" ]; then
  python bench.py --workload volume --depth 9 > gpurun_out/${TAG}_bench_vol513.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_vol513.json"))
print("vol513 value", d["value"], "ms", d["ms_per_step"], "stages", [round(x * 1e3, 1) for x in (d.get("stage_ms") or {}).get("launch_ms", [])], "single", (d.get("single_call") or {}).get("ms"))
PY
  for dd in 8 10; do python profiles/micro/volume_bench.py $dd; done 2>&1 | tee gpurun_out/${TAG}_vol.txt
  ncu --set full --clock-control none --import-source on -k regex:k_vol_rows -s 2 -c 1 -f -o gpurun_out/${TAG}_vol python bench.py --workload volume --depth 9 --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_vol.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_vol.log
fi
if [ "$2" = "launches" ] || [ "$3" = "launches" ] || [ "$4" = "launches" ]; then
  # launch list of the bench command (cold-cache, serialised per-launch times: shares, not absolutes)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-strong > gpurun_out/${TAG}_launches.log 2>&1
  tail -1 gpurun_out/${TAG}_launches.log
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-300
fi
