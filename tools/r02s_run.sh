cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1; tail -4 gpurun_out/r02s_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --ncu > gpurun_out/r02s_bench_s3.json 2> gpurun_out/r02s_bench_s3.err
python -c "
import json;d=json.load(open('gpurun_out/r02s_bench_s3.json'));print('S3', round(d['value']), d['roofline'], 'e2e', d['e2e'], d['e2e_dets_only'], d['sustained'], d['clocks'])" || tail -5 gpurun_out/r02s_bench_s3.err
timeout 300 python bench.py --workload S5 --steps 5 --warmup 3 > gpurun_out/r02s_bench_s5.json 2> gpurun_out/r02s_bench_s5.err
python -c "
import json;d=json.load(open('gpurun_out/r02s_bench_s5.json'));print('S5', round(d['value']), d['roofline']['stage_us_per_cpi'], d['hbm_frac_chain'])" || tail -5 gpurun_out/r02s_bench_s5.err
timeout 300 python bench.py --workload S1 --steps 10 --warmup 3 > gpurun_out/r02s_bench_s1.json 2> gpurun_out/r02s_bench_s1.err
python -c "
import json;d=json.load(open('gpurun_out/r02s_bench_s1.json'));print('S1', d['value'], d['ms_per_step'], d.get('e2e_rows'), d.get('parity'))" || tail -5 gpurun_out/r02s_bench_s1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02s_bench_ref.json 2> gpurun_out/r02s_bench_ref.err; head -c 400 gpurun_out/r02s_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustain 0 > gpurun_out/r02s_l.log 2>&1
