set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_b1fused.py -x -q -m gpu > $O/r3j_pytest.log 2>&1; tail -5 $O/r3j_pytest.log | cut -c1-300
for f in 1 32; do SDG_FUSE_B1=$f timeout 300 python tools/bench_arch.py --arch sngan64 --n 8192 >> $O/r3j_arch.log 2>&1; done
timeout 300 python tools/bench_arch.py --arch sngan32 --n 50000 >> $O/r3j_arch.log 2>&1
cat $O/r3j_arch.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -c 200 --csv --log-file $O/r3j_launches_sngan64.csv python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > $O/r3j_sngan64.log 2>&1
