OUT=gpurun_out; TAG=${1:-r2i}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
for m in 0 7 15 3; do ZS_PDL=$m timeout 120 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err; done
ZS_PDL=7 ZS_LATENT_FWD_ROWS=0 timeout 120 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err
cat $OUT/${TAG}_breakdown.jsonl; tail -3 $OUT/${TAG}_breakdown.err
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "latent and not full_size" > $OUT/${TAG}_sanitizer.log 2>&1; tail -5 $OUT/${TAG}_sanitizer.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_latent_fwd_rows' -s 3 -c 1 -o $OUT/${TAG}_fwd -f python bench.py --steps 2 --warmup 3 --graph 0 --no-e2e --no-cpu-baseline --no-secondary --no-strong > /dev/null 2>&1
