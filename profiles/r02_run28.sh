#!/bin/bash
# final rebuilt library: the transform's dispatch + smoke
timeout 120 python -m pytest tests/test_linear_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
