cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02u_tests.log 2>&1; tail -3 gpurun_out/r02u_tests.log
for i in 1 2; do
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --sustain 0 > gpurun_out/r02u.json 2> gpurun_out/r02u.err
python -c "
import json;d=json.load(open('gpurun_out/r02u.json'));print('default', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_us_per_cpi'].items()}, d['parity']['rdm_rel_err'], d['parity']['flags_differ_unexcused'], d['clocks']['sm_mhz'])" || tail -3 gpurun_out/r02u.err
done
