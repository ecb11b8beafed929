timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_stream -c 1 -f -o gpurun_out/r02_ncu_stream3_d3_4096_wpc1 python scripts/cmp_cfg5.py 3 4096 --wpc=1 --check=0 > gpurun_out/r2i_ncu4096.log 2>&1
tail -2 gpurun_out/r2i_ncu4096.log
