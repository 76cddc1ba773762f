#!/bin/bash
# round 2, call 61 (1 GPU, the last seconds of the budget): smoke() at HEAD (the training defaults changed after call 57)
set -x
mkdir -p gpurun_out/r2
timeout 45 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/c61_smoke.log 2>&1; tail -1 gpurun_out/r2/c61_smoke.log | cut -c1-300
