# ncu --set full captures of the dominant kernels (run under gpurun, 1 GPU); outputs gpurun_out/ncu_<TAG>.ncu-rep.
# SKIP = launch position among the launches matched by KREGEX in one pass of tools/profile_kernels.py (forward) —
# summarise with tools/ncu_summary.py / tools/ncu_hot.py.  Do not put '<' in KREGEX (the shell eats it).
TAG=social_fc      SKIP=3 KREGEX=social_fc_tc_kernel bash tools/ncu_one.sh     # 4th Decoder-2 step's fused social pooling + fc
TAG=gru_dec1       SKIP=0 KREGEX=gru_tc_kernel       bash tools/ncu_one.sh     # Decoder-1 recurrence (12 steps)
TAG=gru_dec2       SKIP=5 KREGEX=gru_tc_kernel       bash tools/ncu_one.sh     # one Decoder-2 step
TAG=deconv2        SKIP=0 KREGEX=deconv_tc_kernel    bash tools/ncu_one.sh     # fused 4x4x128 -> 8x8x64
TAG=deconv3        SKIP=1 KREGEX=deconv_tc_kernel    bash tools/ncu_one.sh     # fused 8x8x64 -> 16x16x32
# large-scene social pooling (cfg3 / cfg5 shapes)
ncu --set full --clock-control none --import-source on -k regex:social_pool_rows_kernel --launch-skip 2 -c 1 \
    -o gpurun_out/ncu_pool_cfg3 python tools/profile_kernels.py --passes 1 --scenes 64 --agents 256 --hidden 256 > gpurun_out/ncu_pool_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:social_pool_rows_kernel --launch-skip 2 -c 1 \
    -o gpurun_out/ncu_pool_cfg5 python tools/profile_kernels.py --passes 1 --scenes 1 --agents 1024 --samples 50 --pred-length 40 --scene-size 512 > gpurun_out/ncu_pool_cfg5.log 2>&1
# launch lists (gpu__time_duration only): forward step and train step
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_fwd.csv python tools/profile_kernels.py --passes 1 --ioc-iters 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py > /dev/null 2>&1
ls -la gpurun_out/
