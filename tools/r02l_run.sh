set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcw_kernel -s 2 -c 1 -o gpurun_out/r02l_pcw -f python bench.py --cpis 16 --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity > gpurun_out/r02l_ncu.log 2>&1
tail -3 gpurun_out/r02l_ncu.log
