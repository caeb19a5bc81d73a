#!/bin/bash
# one short gpurun call: GPU tier + one bench line (no CPU legs)
tag=${1:-q}
shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage5 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench rc $?"
cat gpurun_out/${tag}_gpu_tests.log
grep "kernel" gpurun_out/${tag}_bench.log | head -24
python - <<P
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "drains", d["e2e"].get("stream_drains_per_step"))
P
