import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"],2), round(d["ms_per_step"],4), round(d["e2e"]["value"],2), round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["push_kernel_ms"].items()}, d["raycast_icp"]["raycast_ms"], d["raycast_icp"]["icp_ms"])
    else: print(l.strip()[:300])
