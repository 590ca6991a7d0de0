# A/B: N=16384 transforms as one 512-thread CTA per SM (no register cap) against two of 64 registers (INTT spills 14 words)
for v in "" "FHE_B200_LIB=fhe_study_b200/variants/lib_ntt512.so" ""; do
  echo "== $v" >> gpurun_out/f43_ab.log
  env $v timeout 200 python tools/ntt_ab.py 14 2>&1 | grep -v 4611686 >> gpurun_out/f43_ab.log
done
