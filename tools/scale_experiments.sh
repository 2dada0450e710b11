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
print(json.dumps({'label': '$label', 'n': d['n_gpus'], 'ms': d['ms_per_step'], 'value': d['value'], 'e2e_ms': d['e2e']['ms_per_step'], 'e2e': d['e2e']['value'], 'exchange': d['config'].get('gradient_exchange')}))" >> $OUT
  tail -1 $OUT
}
: > $OUT
if [ -n "$DIAG" ]; then
  run "p2p without pushes" VQA_P2P=1 VQA_P2P_DIAG=nopush
  run "p2p without pushes, barriers" VQA_P2P=1 VQA_P2P_DIAG=nopush,nobarrier
  run "p2p without pushes, barriers, kernel" VQA_P2P=1 VQA_P2P_DIAG=nopush,nokernel
elif [ -n "$ONLY_P2P" ]; then
  run "peer-memory fused reduce-scatter + Adam + all-gather" VQA_P2P=1
else
  run "peer-memory fused reduce-scatter + Adam + all-gather" VQA_P2P=1
  run "NCCL all-reduce, late start" VQA_P2P=0 VQA_EARLY_READY=0
  run "NCCL all-reduce, early start" VQA_P2P=0 VQA_EARLY_READY=1
fi
