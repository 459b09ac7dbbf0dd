#!/bin/bash
# round 2, GPU run 3: full GPU suite on the new kernels (pipelined dot-mode GAT, register-cached softmax, >64 segments,
# wide linear), kernel-variant sweep, bench in both aggregation modes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_generated_models_gpu.py > gpurun_out/r02_pytest3.txt 2>&1
python profiles/variant_bench.py > gpurun_out/r02_variants3.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench3_folded.json 2> gpurun_out/r02_bench3_folded.err
python bench.py --steps 20 --warmup 5 --mode folded_dot --no-cpu-baseline > gpurun_out/r02_bench3_folded_dot.json 2> gpurun_out/r02_bench3_folded_dot.err
tail -5 gpurun_out/r02_pytest3.txt; cat gpurun_out/r02_variants3.txt
python - <<'PY'
import json
for f in ("folded", "folded_dot"):
    try:
        d = json.load(open(f"gpurun_out/r02_bench3_{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["kernel_ms"], {k: v["ms"] for k, v in d.get("kernels", {}).items()})
    except Exception as ex:
        print(f, "FAILED", ex)
PY
