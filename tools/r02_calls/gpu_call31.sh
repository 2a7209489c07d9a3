#!/bin/bash
# the committed last state once more: every GPU test + smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/c31_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c31_pytest.txt
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c31_pytest.txt | tail -8 | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
