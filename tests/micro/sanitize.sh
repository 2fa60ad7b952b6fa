# compute-sanitizer over the small end-to-end paths (memcheck: out-of-bounds / misaligned accesses in any kernel)
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120"
$CS python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
$CS python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "localize or preprocessing_on_device or standalone_pair or raycast_compacting or interpolate_edge" 2>&1 | tail -6
$CS python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k "library_sharded and tiny" 2>&1 | tail -4
