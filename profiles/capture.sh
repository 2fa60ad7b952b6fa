#!/bin/bash
# Everything the summaries under profiles/ are made from.  Run on a GPU box from the repository root:
#   gpurun --timeout 2400 -- 'bash profiles/capture.sh r02'
# then, here:  python profiles/summarize.py r02 C3
# (a number printed by a run under ncu is never a bench value; the launch list's per-launch times are cold-cache and
#  serialised: shares, not absolutes)
R=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list of the whole bench
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep > $O/${R}_launches.log 2>&1
# 2. full sets: the push kernels in the steady state of the bench workload (C3, both lasers per launch), ...
$NCU --set full --import-source on -k regex:'^k_update|^k_classify' --launch-skip 60 -c 6 -o $O/${R}_full_update \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-sweep > $O/${R}_full_update.log 2>&1
# ... the localisation kernels on the same map, ...
$NCU --set full --import-source on -k regex:'^k_raycast|^k_icp' --launch-skip 6 -c 6 -o $O/${R}_full_loc \
    python tests/gpu_perf_icp.py C3 > $O/${R}_full_loc.log 2>&1
# ... and the three scorers (2 x 10^4 hypotheses)
$NCU --set full --import-source on -k regex:'^k_score' --launch-skip 3 -c 6 -o $O/${R}_full_match \
    python tests/gpu_perf_match.py 20000 > $O/${R}_full_match.log 2>&1
ls -la $O/${R}_*
# (the two captures above stop before k_icp and k_score_pdf get their turn: those separately)
$NCU --set full --import-source on -k regex:'^k_icp' --launch-skip 3 -c 3 -o $O/${R}_full_icp python tests/gpu_perf_icp.py C3 > $O/${R}_full_icp.log 2>&1
$NCU --set full --import-source on -k regex:'^k_score_pdf' --launch-skip 1 -c 2 -o $O/${R}_full_pdf python tests/gpu_perf_match.py 20000 > $O/${R}_full_pdf.log 2>&1
