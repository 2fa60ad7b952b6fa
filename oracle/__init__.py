"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the checkers for the CUDA hot path:
  * oracle/_ref/libohm_ref.so  -- the reference's own sources compiled unmodified (oracle/Makefile, `make ref`)
  * oracle/_ref/libohm_port.so -- a plain-C restatement of the same algorithms (oracle/port/, `make port`)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this package.
The product (ohm_tsd_slam_b200) never does.
"""
