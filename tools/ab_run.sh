python -m pytest tests/test_gpu_torus.py tests/test_gpu_chain.py tests/test_gpu_ntt.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -5
python tools/xp_ab.py 4144 32
python tools/xp_ab.py 1184 0
python tools/xp_ab.py 1024 0
python tools/xp_ab.py 148 0
FHE_XP_CT=512 python tools/xp_ab.py 148 0
python tools/xp_ab.py 296 0
FHE_XP_CT=256 python tools/xp_ab.py 296 0
python tools/ntt_ab.py
