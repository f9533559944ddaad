#!/bin/bash
# One GPU-box session: parity tests, the bench line (both arms), the ncu launch list of the same command and one full
# capture of the three hot kernels.  Usage (from the repo root, on the box): bash tools/gpu_session.sh TAG
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
QUICK="--no-e2e --no-cpu-baseline --no-secondary --no-strong"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python tools/parity_budget.py > $OUT/${TAG}_parity.log 2>&1; echo "parity rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --workload vimco --steps 2000 --warmup 200 $QUICK > $OUT/${TAG}_bench_vimco.json 2> $OUT/${TAG}_bench_vimco.err; echo "vimco rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph 0 $QUICK > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_latent_fwd|k_latent_bwd|k_iw_bernoulli_boxf|k_scale_inplace' \
    -s 8 -c 8 -o $OUT/${TAG}_hot -f python bench.py --steps 2 --warmup 3 --graph 0 $QUICK > $OUT/${TAG}_ncu_full.log 2>&1
python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
ls -la $OUT | grep ${TAG}_
