#!/bin/bash
# final state of round 2: every GPU test, the full bench line, launch list, --set full of the recurrent kernel, sanitizer on the
# kernels that changed in session 3 (Xin prefetch: all recurrent instances; K-split GEMM: S = 512), parity report
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/c25_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.txt
grep -E "\[parity\]|passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c25_pytest.txt | tail -30
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c25_bench.txt 2> gpurun_out/c25_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/c25_bench.err
python - <<'P'
import json
for l in open('gpurun_out/c25_bench.txt'):
    if l.startswith('{'):
        d=json.loads(l)
        print('main', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'], 'frac', round(d['roofline']['frac'],4))
        for e in d.get('extra', []):
            print(' ', e['config'][:70], '| value', round(e['value']/1e6,1), 'M/s', '| e2e', round(e.get('e2e',{}).get('value',0)/1e6,1), '| ms', round(e.get('ms_per_step', e.get('ms_per_pass',0)),1), '| frac', round(e.get('roofline_frac',0),4), e.get('clocks',{}).get('sm_mhz'))
        print('cpu', d.get('cpu_baseline'))
P
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/c25_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1; cat gpurun_out/r02_launches_summary.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rnn_tc_kernel -s 7 -c 1 -o gpurun_out/r02_rnn_tc python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/c25_ncu_rnn.log 2>&1
ls -la gpurun_out/*.ncu-rep
MODELS="r941_native_gru r103_native" TOOLS="memcheck racecheck" bash tools/sanitize.sh gpurun_out/sanitize_final
head -3 gpurun_out/sanitize_final/*.log
timeout -s KILL 600 python tools/parity_report.py 64 > gpurun_out/r02_parity_report.txt 2>&1; cat gpurun_out/r02_parity_report.txt | cut -c1-220
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
