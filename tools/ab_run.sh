P="--set full --clock-control none --import-source on"
cap() {  # name kernel-regex skip -- command...
    name=$1; rx=$2; skip=$3; shift 3
    "$@" > gpurun_out/plain_$name.log 2>&1 && ncu $P -k regex:$rx -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name rc=$?"
}
cap s3_xp_pair extprod_fused 2 python tools/prof.py extprod 1024 1 1184
cap s3_l64_n1024 ntt_kernel 3 python tools/prof.py polymul 10 0x3FFFFFFFFFFF0001 65536 4
cap s3_l64_n4096 ntt_kernel 3 python tools/prof.py polymul 12 0x3FFFFFFFFFFF0001 16384 4
python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/s3_bench_ne.json 2>gpurun_out/s3_bench_ne.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/s3_ncu_bench.log 2>&1
echo launches rc=$?
