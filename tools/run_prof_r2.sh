# ncu --set full captures of the round-2 kernels, at most four per gpurun call (the reports are ~12 MB each and
# gpurun_out/ is merged back only below 64 MiB):  bash tools/run_prof_r2.sh A|B
P="--set full --clock-control none --import-source on"
cap() {  # name kernel-regex skip -- command...
    name=$1; rx=$2; skip=$3; shift 3
    "$@" > gpurun_out/plain_$name.log 2>&1 && ncu $P -k regex:$rx -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name rc=$?"
}
if [ "$1" = "A" ]; then
    cap r2_pm1k_u64 ntt_kernel 3 python tools/prof.py polymul 10 65537 65536 4
    cap r2_pm1k_u32 ntt_kernel 3 python tools/prof.py polymul32 10 65537 65536 4
    cap r2_bfv bfv_mul_kernel 2 python tools/prof.py bfv 1048576
    cap r2_kstc ks_tc_kernel 2 python tools/prof.py bootstrap 8192
else
    cap r2_pm16k_gpark ntt_kernel 3 python tools/prof.py polymul 14 65537 4096 4
    cap r2_l64_n1024_b ntt_kernel 3 python tools/prof.py polymul 10 0x3FFFFFFFFFFF0001 65536 4
    cap r2_extprod extprod_fused 2 python tools/prof.py extprod 1024 1 512
fi
