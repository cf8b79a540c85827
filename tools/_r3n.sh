cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_b1fused.py -x -q -m gpu > $O/r3n_pytest.log 2>&1; tail -3 $O/r3n_pytest.log | cut -c1-300
timeout 300 python tools/bench_arch.py --arch sngan64 --n 8192 --iters 10 > $O/r3n_arch.log 2>&1
cat $O/r3n_arch.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:b1_fused -c 6 --csv --log-file $O/r3n_b1.csv python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > $O/r3n_sngan64.log 2>&1
grep b1_fused $O/r3n_b1.csv | tail -2 | cut -c1-50,150-400
