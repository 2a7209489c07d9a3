#!/bin/bash
# 2-GPU sanity of the final state: bench under torchrun (weak scaling, head gate active in every rank), CLI --devices all
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-extra > gpurun_out/c26_bench_2gpu.txt 2> gpurun_out/c26_bench_2gpu.err; echo "rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/c26_bench_2gpu.txt'):
    if l.startswith('{'):
        d=json.loads(l); print('N=2 value', round(d['value']/1e6,1), 'M/s  ms', round(d['ms_per_step'],2), ' e2e', round(d['e2e']['value']/1e6,1), 'M/s', d['clocks'])
P
tail -c 300 gpurun_out/c26_bench_2gpu.err
timeout -s KILL 300 python -m pytest tests/test_host_cli.py -m gpu -q -k "sharded" 2>&1 | tail -2
timeout -s KILL 400 python tools/cli_bench.py 32768 > gpurun_out/c26_cli.txt 2>&1; grep -E "run|identical" gpurun_out/c26_cli.txt | cut -c1-200
