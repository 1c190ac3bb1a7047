#!/bin/bash
# end-of-round check on one GPU: the whole GPU test suite, the smoke entry point, then the evidence refresh (profile_round2b.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_final_tests.log 2>&1; tail -5 gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1; tail -3 gpurun_out/r02_final_smoke.log
bash scripts/profile_round2b.sh train
