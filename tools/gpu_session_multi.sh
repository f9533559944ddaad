#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the 2-rank correctness test, the data-parallel bench at N ranks, and the
# N = 1 diagnosis of the per-step overhead an initialised NCCL communicator adds.
# Usage: bash tools/gpu_session_multi.sh TAG N
TAG=${1:-r2n}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_round2.py -m gpu -q -k "two_ranks or non_current_device" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    tools/allreduce_bench.py > $OUT/${TAG}_allreduce_n$N.jsonl 2> $OUT/${TAG}_allreduce_n$N.err; echo "allreduce bench rc=$?"; cat $OUT/${TAG}_allreduce_n$N.jsonl
QUICK="--no-e2e --no-secondary --no-cpu-baseline"
for n in $(seq 2 $N | awk -v N=$N '$1==2||$1==4||$1==8'); do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $n --steps 1000 --warmup 100 $QUICK > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err; echo "bench n=$n rc=$?"
done
if [ "${3:-}" != "diag" ]; then ls -la $OUT | grep ${TAG}; exit 0; fi
# N = 1 with and without a 1-rank NCCL communicator: step time and per-kernel durations
timeout 300 python bench.py --steps 1000 --warmup 100 $QUICK --no-strong > $OUT/${TAG}_n1_nopg.json 2> $OUT/${TAG}_n1_nopg.err
ZS_BENCH_FORCE_PG=1 timeout 300 python bench.py --steps 1000 --warmup 100 $QUICK --no-strong > $OUT/${TAG}_n1_pg.json 2> $OUT/${TAG}_n1_pg.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_n1_nopg_launches.csv \
    python bench.py --steps 20 --warmup 5 --graph 0 $QUICK --no-strong > /dev/null 2>&1
ZS_BENCH_FORCE_PG=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_n1_pg_launches.csv \
    python bench.py --steps 20 --warmup 5 --graph 0 $QUICK --no-strong > /dev/null 2>&1
ls -la $OUT | grep ${TAG}
