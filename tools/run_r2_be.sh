#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2be_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2be_smoke.log 2>&1
cat gpurun_out/r2be_pytest.log gpurun_out/r2be_smoke.log
