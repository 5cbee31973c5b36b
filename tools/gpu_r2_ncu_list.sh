#!/bin/bash
# ncu launch list of one 128-chunk step (cold-cache, serialised: compare SHARES) — by kernel and by (kernel, grid)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/launches_b128.csv \
    python tools/ncu_workload.py --batch 128 --steps 2 > $O/ncu_workload.log 2>&1
tail -3 $O/ncu_workload.log
python tools/launch_summary.py $O/launches_b128.csv > $O/launches_b128.md; head -30 $O/launches_b128.md
python tools/launch_summary.py $O/launches_b128.csv --by-grid > $O/launches_b128_grid.md; head -50 $O/launches_b128_grid.md
gzip -f $O/launches_b128.csv
