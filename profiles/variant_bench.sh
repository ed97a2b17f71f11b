#!/bin/bash
# developer aid: A/B variant builds of libsandcrate.so (sand_crate_b200/_variants/lib*.so) through bench.py
for v in "$@"; do
  SC_LIB=$PWD/sand_crate_b200/_variants/lib$v.so python bench.py --no-cpu-baseline --steps 100 --warmup 50 --e2e-steps 3 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var_$v.json"))
print("$v", round(d["value"]/1e9,3), round(d["ms_per_step"]*1e3,1), {k:round(x["ms"]*1e3,1) for k,x in d["kernels"].items()})
PY
done
