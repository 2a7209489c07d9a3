#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.txt
grep -E "\[parity\]|passed|failed|^FAILED|^ERROR" gpurun_out/c8_pytest.txt | tail -30
timeout 900 python tools/slots_ab.py > gpurun_out/r02_slots_ab2.txt 2>&1; cat gpurun_out/r02_slots_ab2.txt | tail -40
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c8_bench.txt 2> gpurun_out/c8_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c8_bench.err
python - <<'P'
import json
for l in open('gpurun_out/c8_bench.txt'):
    if l.startswith('{'):
        d=json.loads(l)
        print('main', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'], 'frac', round(d['roofline']['frac'],4), round(d['roofline']['frac_sW_only'],4), d['roofline']['step_breakdown_ms'])
        for e in d.get('extra', []):
            print(' ', e['config'][:70], '| value', round(e['value']/1e6,1), 'M/s', '| e2e', round(e.get('e2e',{}).get('value',0)/1e6,1), '| ms', round(e.get('ms_per_step', e.get('ms_per_pass',0)),1), '| frac', round(e.get('roofline_frac',0),4), e.get('clocks',{}).get('sm_mhz'), e.get('clocks',{}).get('power_w'))
        print('cpu', d.get('cpu_baseline'))
P
tools/sanitize.sh gpurun_out/san2 > /dev/null 2>&1; for f in gpurun_out/san2/sanitizer_*.log; do echo "$f: $(head -1 $f | cut -c1-120)"; done
