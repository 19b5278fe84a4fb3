#!/bin/bash
mkdir -p gpurun_out
TCGEN_KNOBS="TCGEN=1" timeout 200 python tools/tcgen_check.py 100000 2>&1 | cut -c1-330
TCGEN_SHAPES=0 TCGEN_KNOBS="TCGEN=1" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/tcgen_launches.csv python tools/tcgen_check.py 100000 > gpurun_out/tcgen_ncu.log 2>&1; echo "rc=$?"
grep -E "weights_image|tcgen_contract|finalize" gpurun_out/tcgen_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6
