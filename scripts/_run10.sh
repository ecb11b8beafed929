timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alternative or async or cfg5 or executed_work" 2>&1 | tail -5
timeout 600 python scripts/cmp_cfg5.py 3 512 4096 --wpc=4 --check=2 2>&1 | cut -c1-300
timeout 600 python scripts/cmp_cfg5.py 3 512 4096 --wpc=4 --check=0 --prefetch=0 2>&1 | cut -c1-200
timeout 600 python scripts/cmp_cfg5.py 6 512 --wpc=4 --check=1 2>&1 | cut -c1-300
timeout 600 python scripts/cmp_cfg5.py 3 4096 --wpc=1,2 --check=0 2>&1 | cut -c1-200
