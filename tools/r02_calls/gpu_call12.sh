#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_cli.py tests/test_linkproof.py -m gpu -q > gpurun_out/c12_pytest.txt 2>&1; tail -4 gpurun_out/c12_pytest.txt
timeout 900 python tools/cli_bench.py 65536 > gpurun_out/c12_cli_bench.txt 2>&1; cat gpurun_out/c12_cli_bench.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
