#!/bin/bash
# one gpurun call: GPU tier, smoke, bench over the BASELINE configs (ours) + reference arm; everything lands in gpurun_out/
tag=${1:-r02a}
cfgs=${2:-"1 0 4 3 2"}
python -X faulthandler -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
for c in $cfgs; do
	timeout 900 python -X faulthandler bench.py --config $c --steps 5 --warmup 3 --cpu-seconds 12 > gpurun_out/${tag}_bench_c$c.json 2> gpurun_out/${tag}_bench_c$c.log; echo "bench config $c rc $?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.log; echo "ref rc $?"
cat gpurun_out/${tag}_gpu_tests.log; tail -2 gpurun_out/${tag}_smoke.log
for c in $cfgs; do tail -c 600 gpurun_out/${tag}_bench_c$c.log; done
