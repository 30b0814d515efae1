#!/bin/bash
# A/B of the odd A-A step: x-marching rows (kernel 0) against z-walking CTAs (kernel 4), then ncu --set full
# of two launches of each.  Run under gpurun on ONE GPU; outputs in gpurun_out/.
set -u
export EK_B200_LIB=$PWD/ek-pnp-3d_b200/libek_b200_xcheck.so   # the marching kernels live in the cross-check build
TAG=${1:-r02}
mkdir -p gpurun_out
for co in -1 50 100; do
  EK_DEBUG=1 EK_MARCH_CARVEOUT=$co python tools/quick_bench.py 256 256 256 40 kernel=5 2>&1 | tail -2
done
python tools/quick_bench.py 256 256 256 40 kernel=0 2>&1 | tail -1
CMD="python tools/quick_bench.py 256 256 256 6"
ncu --set full --clock-control none --import-source on -k regex:ek_march_kernel -s 2 -c 2 \
    -o gpurun_out/${TAG}_march -f $CMD kernel=5 > gpurun_out/${TAG}_ncu_march.log 2>&1
echo "march capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ek_step_kernel -s 6 -c 2 \
    -o gpurun_out/${TAG}_zwalk -f $CMD kernel=0 > gpurun_out/${TAG}_ncu_zwalk.log 2>&1
echo "zwalk capture rc=$?"
ls -la gpurun_out/${TAG}_*
