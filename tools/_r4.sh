cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_b1fused.py -x -q -m gpu > $O/r4e_pytest.log 2>&1; tail -3 $O/r4e_pytest.log
timeout 300 python tools/conv_microbench.py --b1fused 1 --n 12504 --fused-only 1 > $O/r4e_micro32.log 2>&1; cat $O/r4e_micro32.log
timeout 300 python tools/conv_microbench.py --b1fused 64 --n 8192 --fused-only 1 > $O/r4e_micro64.log 2>&1; cat $O/r4e_micro64.log
timeout 300 python tools/bench_arch.py --arch sngan32 --n 50000 > $O/r4e_arch.log 2>&1; cat $O/r4e_arch.log
