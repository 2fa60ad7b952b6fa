#!/bin/bash
# Builds tests/cpp/slam_loop.cpp against the REFERENCE's own headers + oracle/_ref/libohm_ref.so (build (A) of its
# header) and regenerates tests/golden/slam_loop_*_mode0.txt.  Needs /root/reference (this container only).
set -e
cd "$(dirname "$0")/../.."
make -C oracle ref > /dev/null
g++ -std=c++17 -O2 -fopenmp -ffp-contract=off -w -I/root/reference/src -Ioracle/shim tests/cpp/slam_loop.cpp \
    -o oracle/_ref/slam_loop_ref -Loracle/_ref -lohm_ref -Wl,-rpath,"$PWD/oracle/_ref"
PHI_MIN=$(python3 -c 'import math; print(repr(-135.0*math.pi/180.0))')
R1=$(python3 -c 'import math; print(repr(math.pi/240.0))')
R2=$(python3 -c 'import math; print(repr(math.pi/720.0))')
OMP_NUM_THREADS=1 oracle/_ref/slam_loop_ref tests/golden/scans_tiny.bin 8 0.025 3 361 $R1 $PHI_MIN 8.0 0.001 2.0 0 | grep -v '^init' > tests/golden/slam_loop_tiny_mode0.txt
OMP_NUM_THREADS=1 oracle/_ref/slam_loop_ref tests/golden/scans_C1.bin 10 0.025 3 1081 $R2 $PHI_MIN 30.0 0.001 2.0 0 | grep -v '^init' > tests/golden/slam_loop_C1_mode0.txt
tail -n 2 tests/golden/slam_loop_tiny_mode0.txt tests/golden/slam_loop_C1_mode0.txt
