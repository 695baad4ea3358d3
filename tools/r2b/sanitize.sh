#!/bin/bash
# compute-sanitizer memcheck + racecheck over one eager batch-4 training step (every kernel of the step, including the
# streaming HiddenLayer GEMMs, the fp32 backward-weights launches and the grid barrier); usage: tools/r2b/sanitize.sh <tag>
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
tag=${1:-r2b}
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/r2/one_step.py 1 4 > gpurun_out/${tag}_sanitizer_memcheck.txt 2>&1
tail -4 gpurun_out/${tag}_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/r2/one_step.py 1 4 > gpurun_out/${tag}_sanitizer_racecheck.txt 2>&1
tail -4 gpurun_out/${tag}_sanitizer_racecheck.txt
