OUT=gpurun_out; TAG=${1:-r2g}
mkdir -p $OUT
QUICK="--no-e2e --no-cpu-baseline --no-secondary --no-strong"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
for pdl in 1 0; do
  ZS_PDL=$pdl timeout 300 python bench.py --steps 2000 --warmup 200 $QUICK > $OUT/${TAG}_bench_pdl$pdl.json 2> $OUT/${TAG}_bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"
  ZS_PDL=$pdl timeout 300 python bench.py --workload vimco --steps 2000 --warmup 200 $QUICK > $OUT/${TAG}_vimco_pdl$pdl.json 2> $OUT/${TAG}_vimco_pdl$pdl.err
done
python - <<'PY'
import json
for f in ("bench_pdl1","bench_pdl0","vimco_pdl1","vimco_pdl0"):
    try:
        b=json.load(open("gpurun_out/%s_%s.json" % ("r2g", f)))
        print(f, "step %.2f us  kernel-seq %.2f us  fused %.2f us  step_frac %.3f launches %s" % (b["ms_per_step"]*1e3, b["kernel_sequence"]["ms_per_step"]*1e3, b["roofline"]["kernel_ms"]*1e3, b["roofline"]["step_frac"], b["launches_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph 0 $QUICK > $OUT/${TAG}_ncu_bench.log 2>&1
ls -la $OUT | grep $TAG
