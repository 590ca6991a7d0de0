P="--set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu $P -k regex:$rx -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"; }
python bench.py > gpurun_out/f27_bench.json 2> gpurun_out/f27_bench.err; echo bench rc=$?
python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/f27_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/f27_launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/f27_ncu.log 2>&1; echo launches rc=$?
cap f27_pm1k_u64 ntt_kernel 3 python tools/prof.py polymul 10 65537 65536 4
cap f27_pm1k_u32 ntt_kernel 3 python tools/prof.py polymul32 10 65537 65536 4
cap f27_pm4k_u64 ntt_kernel 3 python tools/prof.py polymul 12 65537 16384 4
