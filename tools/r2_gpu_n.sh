#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -14 gpurun_out/r2n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
