#!/bin/bash
# one gpurun call: GPU tier, smoke, bench of every BASELINE config (ours) + reference arm, ncu launch list, full ncu capture of the top kernels
tag=${1:-r02}
configs=${2:-"1 0 4 3 2"}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
for c in $configs; do
	timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c$c.json 2> gpurun_out/${tag}_bench_c$c.log; echo "bench config $c rc $?"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.log; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
	python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stage5 > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu launch list rc $?"
ncu --set full --import-source on --clock-control none \
	-k regex:"^(k_graph_build|k_group_partition|k_group_partition_warp|k_hit_cigar|k_pair_probe|k_bridge_dp_warp|k_lb_cov_segments|k_cov_add|k_vote_type2|k_pair_decide|k_update|k_hit_bounds|k_hcst_insert|k_frag_group|k_frag_align|k_cluster_emit)$" \
	--launch-skip 19 --launch-count 19 -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stage5 > gpurun_out/${tag}_ncu2.log 2>&1; echo "ncu full rc $?"
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_raw.csv 2>/dev/null; echo "raw csv rc $?"
# gpurun_out/ comes back only if it stays under 64 MiB: the report itself stays on the box when it is large (the raw page has every counter)
sz=$(stat -c %s gpurun_out/${tag}_prof.ncu-rep 2>/dev/null || echo 0); echo "ncu-rep bytes $sz"
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/${tag}_prof.ncu-rep; echo "ncu-rep removed (too large to pull)"; fi
du -sh gpurun_out
cat gpurun_out/${tag}_gpu_tests.log; tail -1 gpurun_out/${tag}_smoke.log
