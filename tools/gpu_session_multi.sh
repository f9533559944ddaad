#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the 2-rank correctness test (three exchange backends), the exchange microbench
# (both kernels of this library against NCCL, with a correctness check) and the data-parallel bench at N ranks.
# Usage: bash tools/gpu_session_multi.sh TAG N [variants]
#   variants: also time the N-rank step with the exchange / bucket-zeroing placements of profiles/r2_notes.md
TAG=${1:-r2n}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then  # the 2-rank correctness test needs two GPUs; larger boxes only run the timings
timeout 400 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_round2.py -m gpu -q -k "two_ranks or non_current_device" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    tools/allreduce_bench.py > $OUT/${TAG}_allreduce_n$N.jsonl 2> $OUT/${TAG}_allreduce_n$N.err; echo "allreduce bench rc=$?"; grep '^{' $OUT/${TAG}_allreduce_n$N.jsonl
QUICK="--no-e2e --no-secondary --no-cpu-baseline"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 1000 --warmup 100 $QUICK > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
  python - <<PY
import json
for l in open("$OUT/${TAG}_${name}_n$N.json"):
    if l.startswith("{"):
        b=json.loads(l); c=b["comm"] or {}
        print("$name: step %.2f us value %.0f M/s  variant %s tuned %s launches %s strong %s" % (b["ms_per_step"]*1e3, b["value"]/1e6, c.get("peer_variant"), c.get("peer_tuned_us"), b["launches_per_step"], [(x["B_global"], round(x["ms_per_step"]*1e3,2)) for x in b.get("strong_scaling", [])]))
PY
}
run bench A=1
if [ "${3:-}" = "variants" ]; then
  QUICK="$QUICK --no-strong"
  run side ZS_BUCKET_STREAM=side
  run nccl ZS_BENCH_COMM=nccl
  run pdl0 ZS_PDL=0
fi
ls -la $OUT | grep ${TAG}_
