#!/bin/bash
# Kernel experiments: the default bench (device-timed leg only) once per library variant in gpurun_variants/*.so.
#   gpurun -- 'bash tools/variant_bench.sh [bench args]'
mkdir -p gpurun_out
shopt -s nullglob
TAG=$(echo "$*" | tr -c "a-zA-Z0-9\n" "_")
for so in "" gpurun_variants/*.so; do
  name=main; [ -n "$so" ] && name=$(basename $so .so)
  SLAMDUNK_B200_LIB=${so:+$PWD/$so} python bench.py --steps 10 --no-e2e --no-cpu-baseline "$@" > gpurun_out/variant_$name$TAG.json 2> gpurun_out/variant_$name$TAG.err
  python - "$name$TAG" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/variant_%s.json" % sys.argv[1]))
    k = d["roofline"]["kernel_split_last_batch"]
    print("%-10s value %.4g  kernel_ms %.4f  lean %.4f staged %.4f plain %.4f" % (sys.argv[1], d["value"], d["roofline"]["kernel_ms"], k["lean_ms"], k["staged_ms"], k["plain_ms"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
