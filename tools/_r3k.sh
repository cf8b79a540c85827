set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"b1_fused" -s 2 -c 1 -o $O/r3k_b1f64 python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > $O/r3k_ncu.log 2>&1
tail -2 $O/r3k_ncu.log
