OUT=gpurun_out; TAG=${1:-r2m}
mkdir -p $OUT
nproc; uptime
for p in 7 0 7 0; do echo "== ZS_PDL=$p"; ZS_PDL=$p timeout 200 python tools/hoststep_bench.py 2>&1 | tail -5; done > $OUT/${TAG}_hoststep.txt 2>&1
cat $OUT/${TAG}_hoststep.txt; uptime
