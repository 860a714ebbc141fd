#!/bin/bash
# One GPU visit: parity suite (default scatter), A/B of the two scatters, bench line, parity suite
# with the tiled scatter, ncu captures of the three pixel kernels, launch list of the bench.
# Every stage writes into gpurun_out/ as it goes (later stages are optional if time runs out).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 120 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "exit $?" >> $O/pytest_default.log)
tail -4 $O/pytest_default.log
(timeout 150 python tools/ab_scatter.py > $O/ab.log 2>&1; echo "exit $?" >> $O/ab.log)
grep -c '"ok": true' $O/ab.log; grep '"ok": false' $O/ab.log | cut -c1-300; grep -v "^compare" $O/ab.log | tail -14
(timeout 200 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench exit $?")
cut -c1-400 $O/bench_final.json
(FM_SCATTER=tiled timeout 120 python -m pytest tests -m gpu -x -q > $O/pytest_tiled.log 2>&1; echo "exit $?" >> $O/pytest_tiled.log)
tail -3 $O/pytest_tiled.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_moments|k_flow_lean|k_distribute' -s 3 -c 3 -f -o $O/r1_v12_pixel_kernels python tools/prof_step.py 150 360 640 2 > $O/ncu_v12.log 2>&1
tail -2 $O/ncu_v12.log
FM_SCATTER=tiled timeout 100 ncu --set full --clock-control none --import-source on -k regex:'k_distribute' -s 1 -c 1 -f -o $O/r1_v12_tiled python tools/prof_step.py 150 360 640 2 > $O/ncu_v12_tiled.log 2>&1
tail -1 $O/ncu_v12_tiled.log
FM_SCATTER=tiled timeout 90 compute-sanitizer --tool racecheck python tools/prof_step.py 3 72 96 1 > $O/racecheck_tiled.log 2>&1
tail -3 $O/racecheck_tiled.log
FM_BENCH_SKIP_CPU=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r1_v12_launches.csv python bench.py --steps 2 --warmup 1 > $O/launches_bench.log 2>&1
tail -1 $O/launches_bench.log | cut -c1-200
