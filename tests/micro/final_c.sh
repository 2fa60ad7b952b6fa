rm -f gpurun_out/r02_full_*.ncu-rep
timeout 900 bash profiles/capture.sh r02 2>&1 | tail -9
timeout 400 python bench.py > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 200 gpurun_out/r02_bench_e.err
