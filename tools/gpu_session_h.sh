OUT=gpurun_out; TAG=${1:-r2h}
mkdir -p $OUT
for m in 0 15 1 2 4 8 9 13; do ZS_PDL=$m timeout 120 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err; done
ZS_PDL=0 ZS_LATENT_BWD_DEEP=0 timeout 120 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err
ZS_PDL=0 ZS_B=128 timeout 120 python tools/step_breakdown.py >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err
cat $OUT/${TAG}_breakdown.jsonl; tail -3 $OUT/${TAG}_breakdown.err
