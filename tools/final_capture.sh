#!/bin/bash
# one gpurun call: GPU tier, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of the top kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.log; echo bench rc $?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; echo ref rc $?
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv \
	python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stage5 > gpurun_out/ncu_final.log 2>&1; echo ncu1 rc $?
ncu --set full --import-source on --clock-control none \
	-k regex:"k_graph_build|k_group_partition|k_hit_cigar|k_qid_insert|k_bridge_dp|k_covc_emit|k_cov_add|k_vote_type2|k_pair|k_update|k_bundle_bounds|k_hcst_insert|k_frag_group" \
	--launch-skip 13 --launch-count 16 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stage5 > gpurun_out/ncu_final2.log 2>&1; echo ncu2 rc $?
cat gpurun_out/gpu_tests_final.log; tail -1 gpurun_out/smoke_final.log
