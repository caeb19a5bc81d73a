#!/bin/bash
AGPU_PIPE_PROFILE=1 timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --no-stage5 --scale 0.25 > gpurun_out/r02h_c4.json 2> gpurun_out/r02h_c4.log; echo "c4 rc $?"
grep "pipe-profile" gpurun_out/r02h_c4.log
AGPU_PIPE_PROFILE=1 timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --no-stage5 --scale 0.25 --streams 1 --no-prefetch > gpurun_out/r02h_c4s1.json 2> gpurun_out/r02h_c4s1.log; echo "c4 1 stream rc $?"
grep "pipe-profile" gpurun_out/r02h_c4s1.log | head -12
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"], d["e2e"]["h2d_bytes_per_step"])
    except Exception as e: print(f, "no json", e)
PY
