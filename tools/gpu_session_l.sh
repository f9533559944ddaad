OUT=gpurun_out; TAG=${1:-r2l}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "latent" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python tools/profile_api_step.py > $OUT/${TAG}_api_profile.txt 2>&1; tail -8 $OUT/${TAG}_api_profile.txt
timeout 300 python tools/hoststep_bench.py > $OUT/${TAG}_hoststep.txt 2>&1; cat $OUT/${TAG}_hoststep.txt
