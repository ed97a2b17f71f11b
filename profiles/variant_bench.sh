# A/B timing of build variants on one GPU: profiles/variant_bench.sh  (run on the GPU box; variants are built in the
# build container first: see the python lines at the bottom of this file)
set -x
for v in base k5mb5 k5mb5r4 k5r8 k5mb3; do
  lib=sand_crate_b200/libsandcrate_$v.so
  [ "$v" = base ] && lib=sand_crate_b200/libsandcrate.so
  [ -f $lib ] || continue
  SC_LIB=$PWD/$lib python bench.py --steps 150 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/variant_$v.json"))
print("$v", round(d["ms_per_step"] * 1e3, 2), {k: round(x["ms"] * 1e3, 1) for k, x in d["kernels"].items()})
PY
done
# build container:
#   python -c "from sand_crate_b200 import build as b; b.build(True, defines=['SC_K5_TILE_MINBLOCKS=5'], out='sand_crate_b200/libsandcrate_k5mb5.so')"
