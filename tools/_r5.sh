cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_size.py -x -q -m gpu > $O/r5b_pytest.log 2>&1; tail -3 $O/r5b_pytest.log
timeout 300 python tools/step_breakdown.py > $O/r5b_breakdown.log 2>&1; cat $O/r5b_breakdown.log
