cd $GRAFT_REPO_ROOT
for s in 1 2 3; do for c in 16 32 64; do
RB200_SLOTS=$s timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk $c > gpurun_out/r02m.json 2> gpurun_out/r02m.err
python -c "
import json;d=json.load(open('gpurun_out/r02m.json'));print('slots $s chunk $c', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_us_per_cpi'].items()})" || tail -3 gpurun_out/r02m.err
done; done
