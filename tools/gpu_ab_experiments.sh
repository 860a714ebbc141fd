#!/bin/bash
# One GPU visit for the opt-in kernels (round 2 opener): parity + timing of the scatter variants
# (FM_SCATTER=tiled / tiled64) and of the staged flow kernel (FM_FLOW_STAGED=1) against the default
# path, the -m gpu suite under each switch, racecheck of the new shared-memory kernels, one ncu
# capture each.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 240 python tools/ab_scatter.py > $O/ab_scatter.log 2>&1; echo "exit $?" >> $O/ab_scatter.log)
grep -c '"ok": true' $O/ab_scatter.log; grep '"ok": false' $O/ab_scatter.log | cut -c1-400; grep -v "^compare" $O/ab_scatter.log | tail -16
(timeout 180 python tools/ab_flow.py > $O/ab_flow.log 2>&1; echo "exit $?" >> $O/ab_flow.log)
grep -v "^compare" $O/ab_flow.log | tail -6; grep '"ok": false' $O/ab_flow.log | cut -c1-300
(FM_SCATTER=tiled64 FM_FLOW_STAGED=1 timeout 120 python -m pytest tests -m gpu -x -q > $O/pytest_experiments.log 2>&1; echo "exit $?" >> $O/pytest_experiments.log)
tail -3 $O/pytest_experiments.log
FM_SCATTER=tiled64 FM_FLOW_STAGED=1 timeout 120 compute-sanitizer --tool racecheck python tools/prof_step.py 3 136 96 1 > $O/racecheck_experiments.log 2>&1
tail -2 $O/racecheck_experiments.log
FM_SCATTER=tiled64 FM_FLOW_STAGED=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_flow_lean|k_distribute' -s 2 -c 2 -f -o $O/r2_experiments python tools/prof_step.py 150 360 640 2 > $O/ncu_experiments.log 2>&1
tail -2 $O/ncu_experiments.log
