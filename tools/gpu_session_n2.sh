TAG=${1:-r2u}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
QUICK="--no-e2e --no-secondary --no-cpu-baseline --no-strong"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 1000 --warmup 100 $QUICK > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
  python - <<PY
import json
for l in open("$OUT/${TAG}_${name}_n$N.json"):
    if l.startswith("{"):
        b=json.loads(l); c=b["comm"] or {}
        print("$name: step %.2f us value %.0f M/s  variant %s tuned %s launches %s" % (b["ms_per_step"]*1e3, b["value"]/1e6, c.get("peer_variant"), c.get("peer_tuned_us"), b["launches_per_step"]))
PY
}
run side_overlap A=1
run main_overlap ZS_BUCKET_STREAM=main
run main_inline ZS_BUCKET_STREAM=main ZS_BUCKET_ZERO=inline
run side_inline ZS_BUCKET_ZERO=inline
run nccl ZS_BENCH_COMM=nccl
run side_overlap_pdl0 ZS_PDL=0
