set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window_moments or score_vectorised or top_indices or blur or odd_sizes or stats or scores" > $O/r3b_pytest.log 2>&1; tail -5 $O/r3b_pytest.log
timeout 300 python tools/bench_stats.py > $O/r3b_stats_new.log 2>&1
timeout 300 python tools/bench_stats.py --n 50000 > $O/r3b_stats_50k.log 2>&1
cat $O/r3b_stats_new.log $O/r3b_stats_50k.log | cut -c1-250
