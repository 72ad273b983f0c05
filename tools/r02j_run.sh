set -x
cd $GRAFT_REPO_ROOT
export RB200_ONEPASS=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --cpis 16 --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk 4 > gpurun_out/r02j_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -s 4 -c 1 -o gpurun_out/r02j_onepass3 -f python bench.py --cpis 16 --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk 4 > gpurun_out/r02j_ncu.log 2>&1
tail -3 gpurun_out/r02j_ncu.log
