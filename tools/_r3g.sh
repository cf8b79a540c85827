set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "top_indices or scores" > $O/r3g_pytest.log 2>&1; tail -3 $O/r3g_pytest.log
timeout 300 python tools/bench_stats.py > $O/r3g_stats.log 2>&1
cat $O/r3g_stats.log | cut -c1-250
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -c 200 --csv --log-file $O/r3g_launches_sngan64.csv python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > $O/r3g_sngan64.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 300 --csv --log-file $O/r3g_launches_sg2.csv python tools/bench_arch.py --arch stylegan2 --size 256 --n 112 --iters 1 > $O/r3g_sg2.log 2>&1
