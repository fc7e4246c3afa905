# usage: TAG=name SKIP=n KREGEX=regex bash tools/ncu_one.sh   (one ncu --set full capture at the bench workload)
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} --launch-skip ${SKIP} -c 1 \
    -o gpurun_out/ncu_${TAG} python tools/profile_kernels.py --passes 1 > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
