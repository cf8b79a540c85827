cd $GRAFT_REPO_ROOT
O=gpurun_out
for w in 2 3 4 6 8; do echo "WSTAGES=$w"; SDG_B1_WSTAGES=$w timeout 300 python tools/bench_arch.py --arch sngan64 --n 8192 --iters 10; done > $O/r3m_arch.log 2>&1
cat $O/r3m_arch.log
