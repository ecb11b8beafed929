set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alternative or async" > gpurun_out/r2b_tests_stream.txt 2>&1; tail -5 gpurun_out/r2b_tests_stream.txt
timeout 600 python scripts/cmp_cfg5.py 3 512 --wpc=1,2,4 > gpurun_out/r2b_cfg5d3_512.txt 2>&1; cat gpurun_out/r2b_cfg5d3_512.txt
timeout 600 python scripts/cmp_cfg5.py 3 4096 --wpc=1,2,4 > gpurun_out/r2b_cfg5d3_4096.txt 2>&1; cat gpurun_out/r2b_cfg5d3_4096.txt
timeout 600 python scripts/cmp_cfg5.py 6 512 --wpc=2,4 --check=1 > gpurun_out/r2b_cfg5d6_512.txt 2>&1; cat gpurun_out/r2b_cfg5d6_512.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_tests_all.txt 2>&1; tail -15 gpurun_out/r2b_tests_all.txt
