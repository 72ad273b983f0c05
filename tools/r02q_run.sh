cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; tail -4 gpurun_out/r02q_tests.log
timeout 200 python bench.py --workload S1 --steps 10 --warmup 3 > gpurun_out/r02q_bench_s1.json 2> gpurun_out/r02q_bench_s1.err
python -c "
import json;d=json.load(open('gpurun_out/r02q_bench_s1.json'));print('S1', d['value'], d['ms_per_step'], d.get('e2e_rows'), d.get('parity'))" || tail -5 gpurun_out/r02q_bench_s1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02q_s1_launches.csv python bench.py --workload S1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_s1_l.log 2>&1
