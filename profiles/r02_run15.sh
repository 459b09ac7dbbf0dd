#!/bin/bash
# round 2, session 3: MODE_GAT_COL v2 (column ids through the staging line, 3 syncs) vs shuffled ids, SFU exp; bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "col or model_dot" > gpurun_out/r02_col_pytest.txt 2>&1; tail -3 gpurun_out/r02_col_pytest.txt
timeout 600 python profiles/variant_bench.py > gpurun_out/r02_variants_col_v2.txt 2>&1; cat gpurun_out/r02_variants_col_v2.txt | cut -c1-330
timeout 600 python bench.py --mode reflected --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_reflected.json 2> gpurun_out/r02_bench_reflected.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_reflected.json').read().strip().splitlines()[-1])
print(d['value'], d['kernel_ms'], d['e2e']['value'], d['parity_rel_err'], d['roofline']['l2'])
PY
tail -5 gpurun_out/r02_bench_reflected.err
