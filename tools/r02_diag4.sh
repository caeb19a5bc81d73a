#!/bin/bash
python -X faulthandler -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02j_gpu_tests.log
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r02j_c3.json 2> gpurun_out/r02j_c3.log; echo "c3 rc $?"
AGPU_PIPE_PROFILE=1 timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r02j_c4.json 2> gpurun_out/r02j_c4.log; echo "c4 rc $?"
cat gpurun_out/r02j_gpu_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02j_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"], d["e2e"]["sub_batches"], "stage5", d.get("stage5",{}).get("ms"), d.get("stage5",{}).get("cpu_baseline",{}).get("pairs_per_sec"), d.get("stage5",{}).get("pairs_per_sec"), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "no json", e)
PY
grep "pipe-profile" gpurun_out/r02j_c4.log | head -14
for c in 4 3; do grep -v "kernel\|pipe-profile" gpurun_out/r02j_c$c.log | tail -3 | cut -c1-300; done
