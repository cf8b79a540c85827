cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_b1fused.py -x -q -m gpu > $O/r4d_pytest.log 2>&1; tail -3 $O/r4d_pytest.log
timeout 300 python tools/conv_microbench.py --b1fused 64 --n 8192 --fused-only 1 > $O/r4d_micro.log 2>&1; cat $O/r4d_micro.log
timeout 300 python tools/bench_arch.py --arch sngan64 --n 8192 > $O/r4d_arch.log 2>&1; cat $O/r4d_arch.log
