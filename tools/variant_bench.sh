#!/bin/bash
# one gpurun call: the bench line (no CPU legs) for the product build and for every tuning variant under aletsch_b200/variants/
tag=${1:-v}
for so in "" aletsch_b200/variants/*.so; do
	[ -n "$so" ] && [ ! -f "$so" ] && continue
	n=${tag}_$(basename "${so:-base}" .so)
	ALETSCH_GPU_LIB=${so:+$PWD/$so} timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage5 > gpurun_out/$n.json 2> gpurun_out/$n.log; echo "bench [$n] rc $?"
	grep "kernel" gpurun_out/$n.log | grep -E "${VB_KERNELS:-.}" | head -${QC_TOP:-12}
	python - <<P
import json
d = json.loads(open("gpurun_out/$n.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"])
P
done
