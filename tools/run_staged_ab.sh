# A/B of the large-degree polymul variants: default (output-row park) against the TMA-staged persistent kernel
for v in "FHE_NTT_GPARK=0 FHE_NTT_STAGED=2" "FHE_NTT_GPARK=0 FHE_NTT_STAGED=2 FHE_NTT_LOGE=4"; do
  echo "== $v" >> gpurun_out/f29_ab.log
  env $v timeout 200 python tools/staged_check.py 13 14 >> gpurun_out/f29_ab.log 2>&1
  env $v timeout 200 python tools/ntt_ab.py 13 14 >> gpurun_out/f29_ab.log 2>&1
done
