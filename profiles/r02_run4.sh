#!/bin/bash
# round 2, GPU run 4 (1 GPU): suite after the ABI changes (bounds_dev, need masks), fast-exp variant, models rebuilt
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest4.txt 2>&1
python profiles/variant_bench.py > gpurun_out/r02_variants4.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
tail -5 gpurun_out/r02_pytest4.txt; cat gpurun_out/r02_variants4.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench4.json"))
print(d["value"], d["e2e"], d["kernel_ms"], {k: v["ms"] for k, v in d.get("kernels", {}).items()})
PY
