#!/bin/bash
# Round-2 opener: validate and time the kernel variants prepared at the end of round 1 (each is
# env-selected and off by default). Run on the GPU box:
#   gpurun --timeout 600 -- 'bash profiles/try_variants.sh'
# Results land in gpurun_out/variants_*.json / .log.
set -u
mkdir -p gpurun_out
DLB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -x -q -k "variant" > gpurun_out/variants_tests.log 2>&1
tail -3 gpurun_out/variants_tests.log
run() { # tag, env assignment, bench arguments
  env $2 timeout 200 python bench.py $3 --no-cpu-baseline > gpurun_out/variants_$1.json 2> gpurun_out/variants_$1.err
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/variants_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1], "value", round(d["value"], 2), "phases", r.get("all_phases_ms") or r.get("avg_launch_ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run c2_base   "DLB_NONE=0"                    "--config c2 --steps 50 --warmup 3"
run c2_chunk  "DOGLEG_GPU_RANGE_CHUNK=4096"   "--config c2 --steps 50 --warmup 3"
run c2_elem   "DOGLEG_GPU_FRONT_ELEM=1"       "--config c2 --steps 50 --warmup 3"
run c2_both   "DOGLEG_GPU_RANGE_CHUNK=4096 DOGLEG_GPU_FRONT_ELEM=1" "--config c2 --steps 50 --warmup 3"
run c3_base   "DLB_NONE=0"                    "--config c3 --steps 10 --warmup 3"
run c3_ring   "DOGLEG_GPU_BATCHED_RING=1"     "--config c3 --steps 10 --warmup 3"
