OUT=gpurun_out; TAG=${1:-r2j}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -x -q -k "latent or graph or sample" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
for cfg in "7 1 1" "7 0 1" "0 1 1" "0 0 1" "7 1 0"; do set -- $cfg; ZS_PDL=$1 ZS_LATENT_FWD_ROWS=$2 ZS_LATENT_BWD_DEEP=$3 timeout 120 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err; done
cat $OUT/${TAG}_breakdown.jsonl; tail -3 $OUT/${TAG}_breakdown.err
QUICK="--no-e2e --no-cpu-baseline --no-secondary --no-strong"
timeout 300 python bench.py --steps 2000 --warmup 200 $QUICK > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload vimco --steps 2000 --warmup 200 $QUICK > $OUT/${TAG}_vimco.json 2> $OUT/${TAG}_vimco.err
python - <<'PY'
import json
for f in ("bench","vimco"):
    try:
        b=json.load(open("gpurun_out/r2j_%s.json" % f))
        print(f, "step %.2f us  kernel-seq %.2f us  fused %.2f us  step_frac %.3f launches %s" % (b["ms_per_step"]*1e3, b["kernel_sequence"]["ms_per_step"]*1e3, b["roofline"]["kernel_ms"]*1e3, b["roofline"]["step_frac"], b["launches_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
