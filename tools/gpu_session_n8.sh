# usage: bash tools/gpu_session_n8.sh TAG N  -- exchange microbench (both kernels, correctness) + the data-parallel bench at N ranks
TAG=${1:-r2v}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo_n$N.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    tools/allreduce_bench.py > $OUT/${TAG}_allreduce_n$N.jsonl 2> $OUT/${TAG}_allreduce_n$N.err; echo "allreduce bench rc=$?"; grep '^{' $OUT/${TAG}_allreduce_n$N.jsonl
QUICK="--no-e2e --no-secondary --no-cpu-baseline"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 1000 --warmup 100 $QUICK > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
python - <<PY
import json
for l in open("$OUT/${TAG}_bench_n$N.json"):
    if l.startswith("{"):
        b=json.loads(l); c=b["comm"] or {}
        print("step %.2f us value %.0f M/s  variant %s tuned %s launches %s" % (b["ms_per_step"]*1e3, b["value"]/1e6, c.get("peer_variant"), c.get("peer_tuned_us"), b["launches_per_step"]))
        print("by rank", b["config"]["ms_per_step_by_rank"]); print("strong", [(x["B_global"], round(x["ms_per_step"]*1e3,2)) for x in b["strong_scaling"]])
PY
tail -3 $OUT/${TAG}_bench_n$N.err
