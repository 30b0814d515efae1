#!/bin/bash
# Profiling pass of the 1-GPU bench command on a B200 (run under gpurun, ONE GPU):
#   1. the plain command (must exit 0),
#   2. ncu launch list with per-launch device time,
#   3. ncu --set full of three LBM launches (lean even/odd) and of the Poisson z-solve.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.json|csv.
set -u
TAG=${1:-r01b}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --pb-iters 20"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
if [ "${3:-}" != "nolist" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
fi
# LBM launches 5..8 of the run: two even and two odd steps of the lean kernel
ncu --set full --clock-control none --import-source on -k regex:ek_step_kernel -s 4 -c 4 \
    -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture (LBM) rc=$?"
if [ "${2:-}" = "poisson" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_zsolve -s 22 -c 2 \
    -o gpurun_out/${TAG}_zsolve -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "full capture (z-solve) rc=$?"
fi
ls -la gpurun_out/${TAG}_*
