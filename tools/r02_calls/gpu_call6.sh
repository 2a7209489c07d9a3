#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.txt
grep -E "\[parity\]|passed|failed|^FAILED|^ERROR" gpurun_out/c6_pytest.txt | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c6_bench.txt 2> gpurun_out/c6_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c6_bench.err
python - <<'P'
import json
for l in open('gpurun_out/c6_bench.txt'):
    if l.startswith('{'):
        d=json.loads(l)
        print('main', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['e2e'].get('d2h_bytes_per_step'), d['clocks'], 'frac', round(d['roofline']['frac'],4))
        for e in d.get('extra', []):
            print(' ', e['config'][:70], '| value', round(e['value']/1e6,1), 'M/s', '| e2e', round(e.get('e2e',{}).get('value',0)/1e6,1), '| ms', round(e.get('ms_per_step', e.get('ms_per_pass',0)),1), '| frac', round(e.get('roofline_frac',0),4), e.get('clocks',{}).get('sm_mhz'))
        print('cpu', d.get('cpu_baseline'))
P
# launch list of two steps (cold-cache, serialised: shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/c6_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1; cat gpurun_out/r02_launches_summary.txt
# --set full of the recurrent kernel and the streamed GEMM (third launch of each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rnn_tc_kernel -s 7 -c 1 -o gpurun_out/r02_rnn_tc python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/c6_ncu_rnn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ws_kernel -s 7 -c 1 -o gpurun_out/r02_gemm_ws python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/c6_ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
