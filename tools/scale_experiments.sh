#!/bin/bash
# All-reduce overlap experiments (N GPUs of one box): bench.py --quick under different reducer / NCCL settings.
#   tools/scale_experiments.sh N outfile
N=${1:-2}; OUT=${2:-gpurun_out/scale_experiments.jsonl}
run() {  # label, env...
  local label=$1; shift
  echo "== $label" >&2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus $N --steps 30 --warmup 5 --quick 2>> gpurun_out/scale_experiments.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'label': '$label', 'n': d['n_gpus'], 'ms': d['ms_per_step'], 'value': d['value'], 'e2e_ms': d['e2e']['ms_per_step'], 'e2e': d['e2e']['value'], 'allreduce': d['config'].get('allreduce')}))" >> $OUT
  tail -1 $OUT
}
: > $OUT
run "late start (hooks at node return)" VQA_EARLY_READY=0
run "early start" VQA_EARLY_READY=1
run "early start, NCCL_MAX_CTAS=8, 8 SMs reserved" VQA_EARLY_READY=1 NCCL_MAX_CTAS=8 VQA_SM_RESERVE=8
run "early start, NCCL_MAX_CTAS=4, 4 SMs reserved" VQA_EARLY_READY=1 NCCL_MAX_CTAS=4 VQA_SM_RESERVE=4
run "early start, NCCL_MAX_CTAS=16, 16 SMs reserved" VQA_EARLY_READY=1 NCCL_MAX_CTAS=16 VQA_SM_RESERVE=16
run "late start, NCCL_MAX_CTAS=8" VQA_EARLY_READY=0 NCCL_MAX_CTAS=8
