# soak run of the seeded random-shape parity tests (GPU against the oracle): bash tools/run_fuzz_soak.sh [seeds]
FHE_FUZZ_SEEDS=${1:-200} timeout 1500 python -m pytest tests/test_gpu_fuzz.py -q -m gpu -x -p no:cacheprovider > gpurun_out/fuzz_soak.log 2>&1
echo "fuzz soak rc=$? seeds=${1:-200}" >> gpurun_out/fuzz_soak.log
tail -3 gpurun_out/fuzz_soak.log
