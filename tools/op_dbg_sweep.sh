# timing experiments of the single-pass kernel (RB200_OP_DBG ablations x chunk sizes); results of dbg != 0 are wrong by design
for c in 8 32 64; do
for d in 0 7 3; do
  RB200_OP_DBG=$d timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --chunk $c > gpurun_out/dbg_$d.json 2> gpurun_out/dbg_$d.err
  python -c "
import json;d=json.load(open('gpurun_out/dbg_$d.json'));print('chunk $c dbg $d', round(d['value']), d['roofline']['stage_us_per_cpi'])" || tail -2 gpurun_out/dbg_$d.err
done
done
