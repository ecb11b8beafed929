set -x
mkdir -p gpurun_out
# ncu full captures of the dominant kernels (one launch each) + launch lists
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_stream -c 1 -f -o gpurun_out/r02_ncu_full_stream_cfg5d3_4096 python scripts/cmp_cfg5.py 3 4096 --check=0 > gpurun_out/r2j_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_stream -c 1 -f -o gpurun_out/r02_ncu_full_stream_cfg5d3_512 python scripts/cmp_cfg5.py 3 512 --check=0 > gpurun_out/r2j_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_stream -c 1 -f -o gpurun_out/r02_ncu_full_stream_cfg5d6_512 python scripts/cmp_cfg5.py 6 512 --check=0 > gpurun_out/r2j_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_warp_kernel -c 1 -f -o gpurun_out/r02_ncu_full_warp_cfg2 python scripts/run_one.py cfg2 kdline 1 > gpurun_out/r2j_d.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2j_e.log 2>&1
tail -2 gpurun_out/r2j_*.log
