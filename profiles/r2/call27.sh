#!/bin/bash
# round 2, call 27 (8 GPUs): final multi-GPU numbers with PDL + serpentine: cfg2 at N = 8 (PDL on / off), N = 4 on the same box, cfg4 at N = 8
set -x
mkdir -p gpurun_out/r2
run() {  # name, nproc, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29530 + RANDOM % 50)) bench.py --gpus $2 ${@:3} > gpurun_out/r2/c27_bench_$1.json 2> gpurun_out/r2/c27_bench_$1.err
  grep '^{' gpurun_out/r2/c27_bench_$1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['e2e']['value'], d.get('parity'), d.get('roofline',{}).get('frac'))" || tail -5 gpurun_out/r2/c27_bench_$1.err
}
run cfg2_8gpu 8 --steps 20 --warmup 5
ANEMOI_B200_PDL=0 run cfg2_8gpu_pdl0 8 --steps 20 --warmup 5
run cfg2_8gpu_b 8 --steps 20 --warmup 5
run cfg2_4gpu 4 --steps 20 --warmup 5
run cfg4_8gpu 8 --workload cfg4 --steps 5 --warmup 3
