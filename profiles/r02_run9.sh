#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tile -s 3 -c 3 -o gpurun_out/r02_edge_tiles -f python profiles/edge_tiles_ncu.py > gpurun_out/r02_edge_tiles_ncu.log 2>&1
tail -3 gpurun_out/r02_edge_tiles_ncu.log
ncu -i gpurun_out/r02_edge_tiles.ncu-rep --page raw --csv > gpurun_out/r02_edge_tiles_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02_edge_tiles_raw.csv
