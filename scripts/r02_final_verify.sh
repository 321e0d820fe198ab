#!/bin/bash
# end-of-round verification on a 2-GPU box: the whole -m gpu suite (incl. the multi-GPU worker) and smoke()
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
cat gpurun_out/r02_final_pytest.log; tail -2 gpurun_out/r02_final_smoke.log
