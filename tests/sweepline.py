import sys,json
d=json.loads(sys.stdin.readline()); print(round(d["value"],2), round(d["ms_per_step"],4), round(d["e2e"]["value"],2), round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["push_kernel_ms"].items()})
for r in d["large_grid_sweep"]["rows"]: print(r["room_m"], round(r["push_ms"],3), round(r["gcell_updates_per_s"],1), round(r["k_update_frac_of_hbm_peak"],3))
