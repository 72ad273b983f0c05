set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass" > gpurun_out/r02i_tests.log 2>&1; tail -5 gpurun_out/r02i_tests.log
for c in 0 2 4; do
RB200_ONEPASS=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --chunk $c > gpurun_out/r02i_bench_c$c.json 2> gpurun_out/r02i_bench_c$c.err
python -c "
import json;d=json.load(open('gpurun_out/r02i_bench_c$c.json'));print('chunk $c', round(d['value']), d['roofline']['stage_us_per_cpi'], d.get('parity'))" || tail -3 gpurun_out/r02i_bench_c$c.err
done
for d in 1 2 3; do
RB200_OP_DBG=$d RB200_ONEPASS=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk 4 > gpurun_out/r02i_dbg$d.json 2> gpurun_out/r02i_dbg$d.err
python -c "
import json;d=json.load(open('gpurun_out/r02i_dbg$d.json'));print('dbg $d', round(d['value']), d['roofline']['stage_us_per_cpi'])" || tail -3 gpurun_out/r02i_dbg$d.err
done
