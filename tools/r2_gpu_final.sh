#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.log
tail -4 gpurun_out/r2_final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
