#!/bin/bash
mkdir -p gpurun_out
python tools/heat_step.py > gpurun_out/r2l_heat.json 2> gpurun_out/r2l_heat.err; cat gpurun_out/r2l_heat.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_heat_launches.csv python tools/heat_step.py --coarse 7 --levels 5 > /dev/null 2>&1
python tools/launch_breakdown.py gpurun_out/r2l_heat_launches.csv > gpurun_out/r2l_heat_breakdown.txt 2>&1; head -30 gpurun_out/r2l_heat_breakdown.txt
gzip -f gpurun_out/r2l_heat_launches.csv
