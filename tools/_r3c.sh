set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"window_moments_tile|select_fused|score_floor_min|score_clip_kernel" -c 8 -o $O/r3c_stats python tools/bench_stats.py --reps 1 > $O/r3c_ncu.log 2>&1
tail -3 $O/r3c_ncu.log
