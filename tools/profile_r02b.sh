set -x
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02b_bench_small.json 2> gpurun_out/r02b_bench_small.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"loglik_kernel|mas2_kernel" --launch-skip 6 -c 2 -f -o gpurun_out/r02b_prof python bench.py --steps 2 --warmup 3 --no-cpu --no-backward > gpurun_out/r02b_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mas_cluster_kernel" --launch-skip 2 -c 1 -f -o gpurun_out/r02b_prof_cluster python tools/cfg4_mas_once.py > gpurun_out/r02b_ncu_cluster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" --launch-skip 4 -c 1 -f -o gpurun_out/r02b_prof_conv python tools/conv_bench.py > gpurun_out/r02b_ncu_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
