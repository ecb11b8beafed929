timeout 900 ncu --set full --clock-control none --import-source on -k regex:kdline_stream -c 1 -f -o gpurun_out/r02_ncu_stream_d3_512_wpc4 python scripts/cmp_cfg5.py 3 512 --wpc=4 --check=0 > gpurun_out/r2d_ncu512.log 2>&1
tail -3 gpurun_out/r2d_ncu512.log
