# A/B timing of build variants on one GPU (variants are built in the build container into sand_crate_b200/libsandcrate_<v>.so)
# usage: bash profiles/variant_bench.sh "<bench args>" v1 v2 ...
set -x
ARGS="$1"; shift
for v in base "$@"; do
  lib=sand_crate_b200/libsandcrate_$v.so
  [ "$v" = base ] && lib=sand_crate_b200/libsandcrate.so
  [ -f $lib ] || continue
  SC_LIB=$PWD/$lib python bench.py $ARGS --no-cpu-baseline --e2e-steps 1 > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/variant_$v.json"))
print("$v", round(d["ms_per_step"] * 1e3, 2), {k: round(x["ms"] * 1e3, 1) for k, x in d["kernels"].items()})
PY
done
