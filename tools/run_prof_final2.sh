# end-of-round captures of the large-degree and TMA-staged polymul kernels, the 62-bit kernel and the external product
P="--set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu $P -k regex:$rx -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"
  # the reports are ~18 MB each and gpurun_out/ is merged back only below 64 MiB: condense on the box, keep the summary
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.csv > /dev/null 2>&1 && rm -f gpurun_out/$name.ncu-rep; }
cap f30_pm16k_gpark ntt_kernel 3 python tools/prof.py polymul 14 65537 4096 4
export FHE_NTT_GPARK=0 FHE_NTT_STAGED=2
cap f30_pm16k_tma ntt_mul_staged 3 python tools/prof.py polymul 14 65537 4096 4
unset FHE_NTT_GPARK FHE_NTT_STAGED
cap f30_pm8k_gpark ntt_kernel 3 python tools/prof.py polymul 13 65537 8192 4
cap f30_l64_n1024 ntt_kernel 3 python tools/prof.py polymul 10 0x3FFFFFFFFFFF0001 65536 4
