python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -c 300 gpurun_out/r02_bench_d.err
