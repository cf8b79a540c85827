set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window_moments or score_vectorised or top_indices or blur or odd_sizes or stats" > $O/r3a_pytest.log 2>&1; tail -5 $O/r3a_pytest.log
timeout 300 python tools/bench_stats.py > $O/r3a_stats_new.log 2>&1
SDG_MOMENTS_GENERIC=1 timeout 300 python tools/bench_stats.py > $O/r3a_stats_generic.log 2>&1
timeout 300 python tools/bench_stats.py --n 50000 > $O/r3a_stats_50k.log 2>&1
timeout 300 python tools/bench_blur.py > $O/r3a_blur.log 2>&1
for v in 0 1 2; do SDG_BLUR_TMA=$v timeout 300 python tools/bench_arch.py --arch stylegan2 --size 256 --n 448 --batch 4 >> $O/r3a_sg2.log 2>&1; done
cat $O/r3a_stats_new.log $O/r3a_stats_generic.log $O/r3a_blur.log $O/r3a_sg2.log | cut -c1-250
