python - <<'PY'
import fhe_study_b200 as fhe
fhe.set_device(0)
for k in (0,1,6,2): print(k, fhe.int_peak(k)/1e12)
PY
