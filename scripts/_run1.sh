set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc; free -g | head -2
python scripts/h2d_floor.py 512 > gpurun_out/r2a_h2d.txt 2>&1
python scripts/run_one.py cfg5d3 kdline 2 4096 > gpurun_out/r2a_cfg5d3_4096.txt 2>&1
python scripts/run_one.py cfg5d3 kdline 2 512 > gpurun_out/r2a_cfg5d3_512.txt 2>&1
python scripts/run_one.py cfg5d6 kdline 1 512 > gpurun_out/r2a_cfg5d6_512.txt 2>&1
python scripts/run_one.py cfg5d6 kdline 1 2048 > gpurun_out/r2a_cfg5d6_2048.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kdline_warpg -c 1 -f -o gpurun_out/r02_ncu_warpg_cfg5d3_512 python scripts/run_one.py cfg5d3 kdline 1 512 > gpurun_out/r2a_ncu512.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_warpg -c 1 -f -o gpurun_out/r02_ncu_warpg_cfg5d3_4096 python scripts/run_one.py cfg5d3 kdline 1 4096 > gpurun_out/r2a_ncu4096.log 2>&1
cat gpurun_out/r2a_*.txt
