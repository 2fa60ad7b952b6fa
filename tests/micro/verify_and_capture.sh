python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -c 600 gpurun_out/r02_bench_c.err
rm -f gpurun_out/r02_full_*.ncu-rep
bash profiles/capture.sh r02 2>&1 | tail -8
