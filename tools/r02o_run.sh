cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" > gpurun_out/r02o_tests.log 2>&1; tail -3 gpurun_out/r02o_tests.log
RB200_NO_FUSED=1 RB200_MTD_TC=1 timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --sustain 0 --chunk 8 > gpurun_out/r02o_b.json 2> gpurun_out/r02o_b.err
python -c "
import json;d=json.load(open('gpurun_out/r02o_b.json'));print('tcgen05 GEMM (RDM only)', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_us_per_cpi'].items()}, d['parity']['rdm_rel_err'])" || tail -3 gpurun_out/r02o_b.err
RB200_NO_FUSED=1 RB200_MTD_TC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtd64_tc -s 1 -c 1 -o gpurun_out/r02o_tc -f python bench.py --cpis 8 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk 8 > gpurun_out/r02o_ncu_tc.log 2>&1
RB200_NO_FUSED=1 RB200_NO_FUSED_V=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtd_fast -s 1 -c 1 -o gpurun_out/r02o_bf -f python bench.py --cpis 8 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk 8 > gpurun_out/r02o_ncu_bf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtd64_tma -s 1 -c 1 -o gpurun_out/r02o_m64 -f python bench.py --cpis 8 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --sustain 0 --no-parity --chunk 8 > gpurun_out/r02o_ncu_m64.log 2>&1
ls -la gpurun_out/r02o*
