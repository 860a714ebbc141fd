#!/bin/bash
# One GPU visit for the opt-in kernels (round 2 opener): parity + timing of FM_SCATTER=tiled /
# tiled64, FM_MAP=tile2d and FM_ADAM=fast against the default path (tools/ab_scatter.py), the
# -m gpu suite under the new switches, racecheck of the new shared-memory kernel, ncu captures.
# Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 240 python tools/ab_scatter.py > $O/ab_scatter.log 2>&1; echo "exit $?" >> $O/ab_scatter.log)
grep -c '"ok": true' $O/ab_scatter.log; grep '"ok": false' $O/ab_scatter.log | cut -c1-400; grep -v "^compare" $O/ab_scatter.log | tail -16
(FM_SCATTER=tiled64 timeout 120 python -m pytest tests -m gpu -x -q > $O/pytest_tiled64.log 2>&1; echo "exit $?" >> $O/pytest_tiled64.log)
tail -3 $O/pytest_tiled64.log
(FM_MAP=tile2d FM_ADAM=fast timeout 120 python -m pytest tests -m gpu -x -q > $O/pytest_tile2d_fastadam.log 2>&1; echo "exit $?" >> $O/pytest_tile2d_fastadam.log)
tail -3 $O/pytest_tile2d_fastadam.log
FM_SCATTER=tiled64 timeout 120 compute-sanitizer --tool racecheck python tools/prof_step.py 3 136 96 1 > $O/racecheck_tiled64.log 2>&1
tail -2 $O/racecheck_tiled64.log
FM_SCATTER=tiled64 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_distribute' -s 1 -c 1 -f -o $O/r2_tiled64 python tools/prof_step.py 150 360 640 2 > $O/ncu_tiled64.log 2>&1
tail -2 $O/ncu_tiled64.log
FM_MAP=tile2d timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_moments|k_distribute' -s 2 -c 2 -f -o $O/r2_tile2d python tools/prof_step.py 150 360 640 2 > $O/ncu_tile2d.log 2>&1
tail -2 $O/ncu_tile2d.log
