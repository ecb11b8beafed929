timeout 1700 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python scripts/san_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -25 gpurun_out/r02_sanitizer_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/san_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -8 gpurun_out/r02_sanitizer_memcheck.log
