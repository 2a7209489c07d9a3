#!/bin/bash
# last state of round 2 (after the ticket batches of the streamed GEMM): every GPU test + the full bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/c29_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c29_pytest.txt
grep -E "\[parity\]|passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c29_pytest.txt | tail -30 | cut -c1-250
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c29_bench.txt 2> gpurun_out/c29_bench.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/c29_bench.txt'):
    if l.startswith('{'):
        d=json.loads(l)
        print('main', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'], 'frac', round(d['roofline']['frac'],4))
        for e in d.get('extra', []):
            print(' ', e['config'][:70], '| value', round(e['value']/1e6,1), 'M/s', '| e2e', round(e.get('e2e',{}).get('value',0)/1e6,1), '| ms', round(e.get('ms_per_step', e.get('ms_per_pass',0)),1), '| frac', round(e.get('roofline_frac',0),4), e.get('clocks',{}).get('sm_mhz'))
        print('cpu', str(d.get('cpu_baseline'))[:200])
P
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
