# ncu captures of the optimizer and sampler kernels (1 GPU)
ncu --set full --clock-control none --import-source on -k regex:fused_merge_update -s 3 -c 1 -o gpurun_out/prof_adam_r1 -f python tools/bench_ops.py --what adam > gpurun_out/ncu_adam.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_adam_r1.csv python tools/bench_ops.py --what adam > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 8 -c 1 -o gpurun_out/prof_sample_r1 -f python tools/bench_ops.py --what sample > gpurun_out/ncu_sample.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sample_r1.csv python tools/bench_ops.py --what sample > /dev/null 2>&1
ls -la gpurun_out
