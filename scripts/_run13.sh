timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alternative or async or cfg5 or executed_work" 2>&1 | tail -3
timeout 600 python scripts/cmp_cfg5.py 3 512 1024 2048 4096 --wpc=-1 --check=1 2>&1 | cut -c1-130
timeout 600 python scripts/cmp_cfg5.py 6 512 4096 --wpc=-1 --check=1 2>&1 | cut -c1-200
