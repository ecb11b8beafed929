timeout 600 python scripts/cmp_cfg5.py 3 2048 4096 --wpc=2 --check=0 2>&1 | cut -c1-330
timeout 600 python scripts/cmp_cfg5.py 3 4096 --wpc=1 --check=0 2>&1 | cut -c1-330
