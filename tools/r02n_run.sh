cd $GRAFT_REPO_ROOT
for s in 2 3; do for c in 4 8 16; do
RB200_COEXIST=1 RB200_SLOTS=$s timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --chunk $c > gpurun_out/r02n.json 2> gpurun_out/r02n.err
python -c "
import json;d=json.load(open('gpurun_out/r02n.json'));print('coexist slots $s chunk $c', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_us_per_cpi'].items()}, d['parity']['rdm_rel_err'], d['parity']['flags_differ_unexcused'])" || tail -3 gpurun_out/r02n.err
done; done
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity > gpurun_out/r02n.json 2> gpurun_out/r02n.err
python -c "
import json;d=json.load(open('gpurun_out/r02n.json'));print('default', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_us_per_cpi'].items()})" || tail -3 gpurun_out/r02n.err
