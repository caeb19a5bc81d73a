#!/bin/bash
# one short gpurun call: GPU tier + one bench line (no CPU legs); further arguments: NAME=VALUE environment variants to bench as well
tag=${1:-q}
shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_gpu_tests.log
cat gpurun_out/${tag}_gpu_tests.log
for v in "" "$@"; do
	n=${tag}_bench${v:+_${v//[^A-Za-z0-9]/_}}
	env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage5 > gpurun_out/$n.json 2> gpurun_out/$n.log; echo "bench [$v] rc $?"
	grep "kernel" gpurun_out/$n.log | head -${QC_TOP:-12}
	python - <<P
import json
d = json.loads(open("gpurun_out/$n.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "drains", d["e2e"].get("stream_drains_per_step"))
P
done
