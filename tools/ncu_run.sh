# ncu captures of the dominant kernels at the bench workload (run under gpurun, 1 GPU).
# Kernel indices are launch positions among the kernels matched by -k in one pass of tools/profile_kernels.py.
N="ncu --set full --clock-control none --import-source on"
P="python tools/profile_kernels.py --passes 1"
timeout 200 $N -k regex:gemm_tc_kernel --launch-skip 144 -c 1 -o gpurun_out/ncu_${TAG}_social_fc_gemm $P > gpurun_out/ncu2.log 2>&1
timeout 200 $N -k regex:gemm_tc_kernel --launch-skip 6 -c 1 -o gpurun_out/ncu_${TAG}_deconv3_gemm $P > gpurun_out/ncu3.log 2>&1
timeout 200 $N -k regex:colbn_act_v4 --launch-skip 5 -c 1 -o gpurun_out/ncu_${TAG}_col2im_d3 $P > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/
