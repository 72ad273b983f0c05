set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_tests.log 2>&1; tail -5 gpurun_out/r02k_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02k_bench.json'));print('pcw', round(d['value']), d['roofline']['stage_us_per_cpi'], d.get('parity'))" || tail -3 gpurun_out/r02k_bench.err
RB200_NO_PCW=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity > gpurun_out/r02k_bench_old.json 2> gpurun_out/r02k_bench_old.err
python -c "
import json;d=json.load(open('gpurun_out/r02k_bench_old.json'));print('old', round(d['value']), d['roofline']['stage_us_per_cpi'])" || tail -3 gpurun_out/r02k_bench_old.err
timeout 200 python bench.py --workload S5 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 > gpurun_out/r02k_bench_s5.json 2> gpurun_out/r02k_bench_s5.err
python -c "
import json;d=json.load(open('gpurun_out/r02k_bench_s5.json'));print('s5', round(d['value']), d['roofline']['stage_us_per_cpi'], d.get('parity'))" || tail -3 gpurun_out/r02k_bench_s5.err
