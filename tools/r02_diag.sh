#!/bin/bash
# diagnostic: cfg1 after the kernel changes; cfg4 end-to-end bisect (upload format x results)
python -X faulthandler -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02e_gpu_tests.log
timeout 600 python bench.py --config 1 --steps 5 --warmup 3 --cpu-seconds 8 --no-stage5 > gpurun_out/r02e_c1.json 2> gpurun_out/r02e_c1.log; echo "c1 rc $?"
for up in compact lean; do for res in counts full; do
	timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --no-stage5 --upload $up --results $res > gpurun_out/r02e_c4_${up}_${res}.json 2> gpurun_out/r02e_c4_${up}_${res}.log; echo "c4 $up $res rc $?"
done; done
AGPU_TILE_MIN_OPS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02e_gpu_tests_tiles.log
cat gpurun_out/r02e_gpu_tests.log gpurun_out/r02e_gpu_tests_tiles.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02e_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"], d["e2e"]["stream_drains_per_step"])
    except Exception as e: print(f, "no json", e)
PY
grep kernel gpurun_out/r02e_c1.log | head -14
