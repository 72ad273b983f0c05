cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "S1 or MTD or mtd or dmx or DMX or windows or mex" > gpurun_out/r02r_tests.log 2>&1; tail -6 gpurun_out/r02r_tests.log
timeout 200 python bench.py --workload S1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_s1.json 2> gpurun_out/r02r_bench_s1.err
python -c "
import json;d=json.load(open('gpurun_out/r02r_bench_s1.json'));print('S1', d['value'], d['ms_per_step'], d.get('e2e_rows'))" || tail -5 gpurun_out/r02r_bench_s1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02r_s1_launches.csv python bench.py --workload S1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_s1_l.log 2>&1
timeout 100 python tools/run_dmx.py > gpurun_out/r02r_dmx.log 2>&1; tail -3 gpurun_out/r02r_dmx.log
