cd $GRAFT_REPO_ROOT
for d in 0 1 0 1; do
RB200_PCW_DBG=$d timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity > gpurun_out/r02t.json 2> gpurun_out/r02t.err
python -c "
import json;d=json.load(open('gpurun_out/r02t.json'));print('dbg $d', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_us_per_cpi'].items()}, d['clocks']['sm_mhz'])" || tail -3 gpurun_out/r02t.err
done
