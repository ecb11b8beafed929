timeout 1700 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 60 python scripts/san_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; grep -E "OK|MISMATCH|SUMMARY|agree" gpurun_out/r02_sanitizer_racecheck.log | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alternative or async or cfg5 or executed_work" 2>&1 | tail -3
timeout 600 python scripts/cmp_cfg5.py 3 512 1024 4096 --wpc=-1 --check=1 2>&1 | cut -c1-130
timeout 600 python scripts/cmp_cfg5.py 6 512 --wpc=-1 --check=1 2>&1 | cut -c1-130
