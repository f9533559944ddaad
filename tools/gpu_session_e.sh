OUT=gpurun_out; TAG=r2e
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/${TAG}_vimco_launches.csv \
    python bench.py --workload vimco --steps 2 --warmup 3 --graph 0 --no-e2e --no-cpu-baseline --no-secondary --no-strong > $OUT/${TAG}_vimco_ncu.log 2>&1
timeout 300 python bench.py --workload vimco --steps 1000 --warmup 100 --no-e2e --no-cpu-baseline --no-secondary --no-strong > $OUT/${TAG}_vimco_bench.json 2>$OUT/${TAG}_vimco_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_latent_fwd_fast' -s 3 -c 1 -o $OUT/${TAG}_fwd -f python bench.py --steps 2 --warmup 3 --graph 0 --no-e2e --no-cpu-baseline --no-secondary --no-strong > /dev/null 2>&1
ls -la $OUT | grep $TAG
