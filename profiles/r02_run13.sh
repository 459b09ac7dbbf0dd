#!/bin/bash
mkdir -p gpurun_out
echo "== contiguous row ranges (current)" > gpurun_out/r02_linear_assignment.txt
timeout 300 python profiles/linear_bench.py >> gpurun_out/r02_linear_assignment.txt 2>&1
echo "== round-robin 128-row tiles (previous)" >> gpurun_out/r02_linear_assignment.txt
GALA_B200_LIB=gala-gnn-acceleration-language_b200/variants/linear_roundrobin.so timeout 300 python profiles/linear_bench.py >> gpurun_out/r02_linear_assignment.txt 2>&1
cat gpurun_out/r02_linear_assignment.txt
( time timeout 900 python bench.py --no-kernels > gpurun_out/r02_bench_generated.json 2> gpurun_out/r02_bench_generated.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_generated.json').read().strip().splitlines()[-1])
print(d['value'], d['kernel_ms'])
for r in d['generated_programs'].get('runs', [d['generated_programs']]): print(r)
PY
