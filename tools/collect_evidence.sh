set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
# A: ncu of one SNGAN-32 sweep (warm-up passes skipped: 2 x 7 tensor-core launches) and the launch list of the bench step
ncu --set full --clock-control none --import-source on -k regex:"b1_fused|conv_swap" -s 14 -c 7 -o $O/r2i_sweep python tools/bench_arch.py --arch sngan32 --n 12504 --iters 1 > $O/r2i_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2j_launches_bench_step.csv python bench.py --steps 2 --warmup 1 --no-eager > $O/r2j_bench_under_ncu.log 2>&1
# B: bench lines
python bench.py > $O/r2d_bench_sngan32.log 2>&1
python bench.py --workload sngan64 > $O/r2d_bench_sngan64.log 2>&1
python bench.py --workload stylegan2 > $O/r2d_bench_stylegan2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2d_bench_reference.log 2>&1
# C: breakdown, loader, architectures, DRS
python tools/step_breakdown.py > $O/r2e_breakdown.log 2>&1
python tools/bench_loader.py > $O/r2e_loader.log 2>&1
for a in sngan32 sngan64 dcgan32; do python tools/bench_arch.py --arch $a --n $([ $a = sngan64 ] && echo 8192 || echo 50000) >> $O/r2e_arch.log 2>&1; done
python tools/bench_arch.py --arch stylegan2 --size 256 --n 512 --batch 4 >> $O/r2e_arch.log 2>&1
python tools/bench_drs.py > $O/r2e_drs.log 2>&1
tail -n 2 $O/r2d_bench_sngan32.log | cut -c1-200; tail -n 3 $O/r2e_arch.log; tail -3 $O/r2e_breakdown.log
