for v in t256pf t256 t384 t512; do
  echo "=== $v"
  FPS_B200_LIB=$PWD/fpsample_b200/variants/libfps_$v.so timeout 600 python scripts/cmp_cfg5.py 3 512 4096 --wpc=1,2,4 --check=1 2>&1 | cut -c1-330
done > gpurun_out/r2c_variants.txt 2>&1
cat gpurun_out/r2c_variants.txt
