#!/bin/bash
# diagnostic run f: tests, cfg1 with all legs, cfg4 end-to-end with the pipeline trace
python -X faulthandler -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02f_gpu_tests.log
timeout 900 python bench.py --config 1 --steps 5 --warmup 3 --cpu-seconds 10 > gpurun_out/r02f_c1.json 2> gpurun_out/r02f_c1.log; echo "c1 rc $?"
AGPU_PIPE_TRACE=1 timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --no-stage5 > gpurun_out/r02f_c4.json 2> gpurun_out/r02f_c4.log; echo "c4 rc $?"
cat gpurun_out/r02f_gpu_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"], d["e2e"]["stream_drains_per_step"], d.get("support"), d.get("stage5",{}).get("ms"))
    except Exception as e: print(f, "no json", e)
PY
grep trace gpurun_out/r02f_c4.log | head -40
grep kernel gpurun_out/r02f_c1.log | head -50
