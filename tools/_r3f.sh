set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "top_indices or scores" > $O/r3f_pytest.log 2>&1; tail -3 $O/r3f_pytest.log
timeout 300 python tools/bench_stats.py > $O/r3f_stats.log 2>&1
timeout 300 python tools/bench_stats.py --n 50000 > $O/r3f_stats_50k.log 2>&1
cat $O/r3f_stats.log $O/r3f_stats_50k.log | cut -c1-250
