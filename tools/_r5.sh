cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_b1fused.py tests/test_gpu_parity.py -x -q -m gpu > $O/r5a_pytest.log 2>&1; tail -3 $O/r5a_pytest.log
timeout 300 python tools/conv_microbench.py --b1fused 64 --n 8192 --fused-only 1 > $O/r5a_micro64.log 2>&1; cat $O/r5a_micro64.log
timeout 300 python tools/bench_arch.py --arch sngan64 --n 8192 > $O/r5a_arch.log 2>&1; cat $O/r5a_arch.log
timeout 300 python tools/step_breakdown.py > $O/r5a_breakdown.log 2>&1; cat $O/r5a_breakdown.log
