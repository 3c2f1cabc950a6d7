import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]
print(round(d["value"],1), 'ms/step', round(d["ms_per_step"],2), 'sweeps', r["sweeps_per_step"], 'ms/sweep', round(r["avg_launch_ms"],3), 'frac', round(r["frac"],3), d["config"]["specialised_sweeps"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
