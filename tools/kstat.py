"""Registers, spills and code size (bytes of SASS) of the NTT kernels in a built object:
  python tools/kstat.py fhe_study_b200/_build/ntt_inst_lazy64.o [filter-regex]"""
import re
import subprocess
import sys

obj = sys.argv[1]
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
sizes = {}
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
for m in re.finditer(r"^\s*[0-9a-f]+\s+[0-9a-f]+\s+([0-9a-f]+)\s.*PROGBITS.*\s\.text\.(\S+)", elf, re.M):
    sizes[m.group(2)] = int(m.group(1), 16)
rows = []
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    name, reg, stack = m.group(1), int(m.group(2)), int(m.group(3))
    k = re.search(r"ntt_kernelINS_\d*([A-Za-z0-9]+)ELi(\d+)ELi(\d+)ELi(\d+)E([mj])", name)
    if not k:
        continue
    label = f"{k.group(1)} logn={int(k.group(2)):2d} loge={k.group(3)} mode={k.group(4)} io={'u64' if k.group(5)=='m' else 'u32'}"
    if flt and not flt.search(label):
        continue
    rows.append((label, reg, stack, sizes.get(name, 0)))
for r in sorted(rows):
    print(f"{r[0]}  reg={r[1]:3d} stack={r[2]:4d} code={r[3] / 1024:6.1f} KB")
