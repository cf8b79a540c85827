set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window_moments or odd_sizes or scores or score_vec" > $O/r3e_pytest.log 2>&1; tail -3 $O/r3e_pytest.log
timeout 300 python tools/bench_stats.py > $O/r3e_stats_new.log 2>&1
SDG_MOMENTS_GENERIC=2 timeout 300 python tools/bench_stats.py > $O/r3e_stats_2buf.log 2>&1
cat $O/r3e_stats_new.log $O/r3e_stats_2buf.log | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"select_fused" -c 1 -o $O/r3e_sel python tools/bench_stats.py --reps 1 > $O/r3e_ncu.log 2>&1
