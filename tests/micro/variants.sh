#!/bin/bash
# exploratory: time k_update variants prebuilt under ohm_tsd_slam_b200/variants/ (see profiles/r02_notes.md)
cd "$(dirname "$0")/../.."
L=ohm_tsd_slam_b200/libtsdslam_b200.so
cp $L /tmp/main.so
for v in "$@"; do
  cp ohm_tsd_slam_b200/variants/$v.so $L
  echo "== $v"
  case $v in
    p*) python tests/gpu_prof_update.py C3 2>&1 | tail -2;;
    *) python tests/gpu_perf_update.py C3 8 quick 2>&1 | grep "batch=2"; python tests/gpu_perf_update.py C2 8 quick 2>&1 | grep "batch=2 all";;
  esac
done
cp /tmp/main.so $L
