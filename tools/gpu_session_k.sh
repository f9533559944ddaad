OUT=gpurun_out; TAG=${1:-r2k}
mkdir -p $OUT
for cfg in "8192 1" "8192 0" "128 1" "128 0" "4096 1" "4096 0"; do set -- $cfg; ZS_B=$1 ZS_LATENT_FWD_ROWS=$2 timeout 200 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err; done
cat $OUT/${TAG}_breakdown.jsonl; tail -3 $OUT/${TAG}_breakdown.err
