# Evidence run of the final build on one B200 (gpurun -- 'bash tools/collect_evidence.sh'); everything lands in gpurun_out/r4_*.
# (r3_* = the same run before the b1_fused_kernel<64> rework; its StyleGAN2 / blur / statistics parts are unchanged and not repeated)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
# A: the GPU suite
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r4_pytest.log 2>&1; tail -4 $O/r4_pytest.log
# B: bench lines
python bench.py > $O/r4_bench_sngan32.log 2>&1
python bench.py --workload sngan64 > $O/r4_bench_sngan64.log 2>&1
python bench.py --workload stylegan2 > $O/r4_bench_stylegan2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/r4_bench_reference.log 2>&1
# C: ncu -- launch lists of the bench step and of one SNGAN-64 pass, --set full of the fused SNGAN-64 block 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r4_launches_bench_step.csv python bench.py --steps 2 --warmup 1 --no-eager > $O/r4_bench_under_ncu.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -c 200 --csv --log-file $O/r4_launches_sngan64.csv python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"b1_fused" -s 2 -c 1 -f -o $O/r4_b1fused64 python tools/bench_arch.py --arch sngan64 --n 8192 --iters 1 > $O/r4_b1fused64_ncu.log 2>&1
# D: breakdowns
python tools/step_breakdown.py > $O/r4_breakdown.log 2>&1
for a in sngan32 sngan64 dcgan32; do python tools/bench_arch.py --arch $a --n $([ $a = sngan64 ] && echo 8192 || echo 50000) >> $O/r4_arch.log 2>&1; done
python tools/conv_microbench.py --b1fused 64 --n 8192 > $O/r4_b1fused64_micro.log 2>&1
tail -n 1 $O/r4_bench_sngan32.log | cut -c1-300; tail -n 1 $O/r4_bench_sngan64.log | cut -c1-300; tail -n 1 $O/r4_bench_stylegan2.log | cut -c1-300; cat $O/r4_arch.log
