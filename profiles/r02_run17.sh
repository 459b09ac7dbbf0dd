#!/bin/bash
# round 2, session 3 (2 GPUs): partitioned forward in the reflected basis -- parity (tests/dist_gpu_check.py), bench N=2, N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_dist_gpu.py -m gpu -x -q -k "col or model_dot or two_gpu" > gpurun_out/r02_col_pytest.txt 2>&1; tail -3 gpurun_out/r02_col_pytest.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_check.py 2>&1 | grep "rel err" | sort > gpurun_out/r02_dist_check_n2.txt; cat gpurun_out/r02_dist_check_n2.txt
for m in reflected folded; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --mode $m --no-cpu-baseline > gpurun_out/r02_bench_n2_$m.json 2> gpurun_out/r02_bench_n2_$m.err; python - $m <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_bench_n2_%s.json'%sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], d['value'], d['kernel_ms'], d['e2e']['value'], d['parity_rel_err'], d['config'].get('mode'))
PY
tail -3 gpurun_out/r02_bench_n2_$m.err
done
timeout 600 python bench.py --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_reflected.json 2> gpurun_out/r02_bench_reflected.err; python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02_bench_reflected.json').read().strip().splitlines()[-1])
print(d['config']['mode'], d['value'], d['kernel_ms'], d['e2e']['value'], d['parity_rel_err'], d['gpu_launches'])
PY
