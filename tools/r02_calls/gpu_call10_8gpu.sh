#!/bin/bash
# 8 GPUs of one box: the bench under torchrun (weak scaling of configs[1] + the strong-scaling configs[3] set), the
# command line files -> fastq at 1 and 8 GPUs, and the multi-device CLI test on real devices
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_8gpu_smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_scale8.txt 2> gpurun_out/r02_scale8.err
echo "bench8 rc=$?"; tail -c 400 gpurun_out/r02_scale8.err
python - <<'P'
import json
for l in open('gpurun_out/r02_scale8.txt'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=8 main', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value']/1e6,1), d['clocks'])
        for e in d.get('extra', []):
            print(' ', e['config'][:80], '| value', round(e['value']/1e6,1), 'M/s | e2e', round(e.get('e2e',{}).get('value',0)/1e6,1), '| ms', round(e.get('ms_per_step', e.get('ms_per_pass',0)),1), e.get('shard_imbalance'))
P
timeout 1200 python tools/cli_bench.py 131072 > gpurun_out/r02_cli_bench.txt 2>&1; cat gpurun_out/r02_cli_bench.txt
timeout 600 python -m pytest tests/test_host_cli.py -m gpu -q -k "devices or trace" > gpurun_out/r02_8gpu_pytest.txt 2>&1; tail -3 gpurun_out/r02_8gpu_pytest.txt
