set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_b1fused.py -x -q -m gpu -k "fused64 or sngan64" -s > $O/r3h_pytest.log 2>&1; tail -40 $O/r3h_pytest.log | cut -c1-300
