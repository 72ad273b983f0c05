cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r02p_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_peak.py --json gpurun_out/r02p_pcie_8.json > gpurun_out/r02p_pcie_8.log 2>&1; tail -4 gpurun_out/r02p_pcie_8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_peak.py --bind --json gpurun_out/r02p_pcie_8_bind.json > gpurun_out/r02p_pcie_8_bind.log 2>&1; tail -4 gpurun_out/r02p_pcie_8_bind.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tools/pcie_peak.py --json gpurun_out/r02p_pcie_4.json > gpurun_out/r02p_pcie_4.log 2>&1; tail -4 gpurun_out/r02p_pcie_4.log
python tools/pcie_peak.py --json gpurun_out/r02p_pcie_1.json > gpurun_out/r02p_pcie_1.log 2>&1; tail -4 gpurun_out/r02p_pcie_1.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err
python -c "
import json;d=json.load(open('gpurun_out/r02p_bench_n8.json'));print('N=8', round(d['value']), 'e2e', d['e2e'], 'dets_only', d.get('e2e_dets_only'), 'gather', d.get('gather_ms'))" || tail -5 gpurun_out/r02p_bench_n8.err
