# usage: bash tools/gpu_session_n.sh TAG N   -- the data-parallel bench at N ranks (quick form)
TAG=${1:-r2t}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
QUICK="--no-e2e --no-secondary --no-cpu-baseline"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 1000 --warmup 100 $QUICK > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
python - <<PY
import json
for l in open("$OUT/${TAG}_bench_n$N.json"):
    if l.startswith("{"):
        b=json.loads(l); c=b["comm"]
        print("step %.2f us value %.0f M/s  variant %s tuned %s launches %s" % (b["ms_per_step"]*1e3, b["value"]/1e6, c.get("peer_variant"), c.get("peer_tuned_us"), b["launches_per_step"]))
        print("strong", [(x["B_global"], round(x["ms_per_step"]*1e3,2)) for x in b["strong_scaling"]])
PY
tail -3 $OUT/${TAG}_bench_n$N.err
