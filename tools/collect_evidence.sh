# Round-2 evidence run on one B200 (gpurun -- 'bash tools/collect_evidence.sh'); everything lands in gpurun_out/r3_*.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
# A: the GPU suite
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r3_pytest.log 2>&1; tail -4 $O/r3_pytest.log
# B: bench lines
python bench.py > $O/r3_bench_sngan32.log 2>&1
python bench.py --workload sngan64 > $O/r3_bench_sngan64.log 2>&1
python bench.py --workload stylegan2 > $O/r3_bench_stylegan2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/r3_bench_reference.log 2>&1
# C: ncu -- launch list of the bench step, --set full of the fused SNGAN-64 block 1 and of one SNGAN-64 sweep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r3_launches_bench_step.csv python bench.py --steps 2 --warmup 1 --no-eager > $O/r3_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"b1_fused|conv_swap|conv_pair" -s 10 -c 10 -o $O/r3_sngan64_sweep python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > $O/r3_sngan64_sweep.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -c 200 --csv --log-file $O/r3_launches_sngan64.csv python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -c 300 --csv --log-file $O/r3_launches_sg2.csv python tools/bench_arch.py --arch stylegan2 --size 256 --n 112 --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blur_tma" -s 6 -c 2 -o $O/r3_blur_tma python tools/bench_arch.py --arch stylegan2 --size 256 --n 112 --iters 1 > /dev/null 2>&1
# D: breakdowns
python tools/step_breakdown.py > $O/r3_breakdown.log 2>&1
python tools/bench_stats.py > $O/r3_stats.log 2>&1
python tools/bench_stats.py --n 50000 >> $O/r3_stats.log 2>&1
python tools/bench_blur.py > $O/r3_blur.log 2>&1
for a in sngan32 sngan64 dcgan32; do python tools/bench_arch.py --arch $a --n $([ $a = sngan64 ] && echo 8192 || echo 50000) >> $O/r3_arch.log 2>&1; done
python tools/bench_arch.py --arch stylegan2 --size 256 --n 512 --batch 4 >> $O/r3_arch.log 2>&1
python tools/conv_microbench.py --b1fused 64 --n 8192 > $O/r3_b1fused64_micro.log 2>&1
tail -n 1 $O/r3_bench_sngan32.log | cut -c1-300; tail -n 1 $O/r3_bench_sngan64.log | cut -c1-300; tail -n 1 $O/r3_bench_stylegan2.log | cut -c1-300; cat $O/r3_arch.log
