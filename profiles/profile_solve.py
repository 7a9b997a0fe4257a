"""Runs a few cfg1-shaped solves for ncu (launch lists / --set full captures).  Not a benchmark:
numbers printed under a profiler are never bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(os.environ.get("PROF_B", "32")); T = int(os.environ.get("PROF_T", "1000")); n = int(os.environ.get("PROF_SOLVES", "1"))
steps = int(os.environ.get("PROF_STEPS", "10"))
dev = torch.device("cuda:0")
m = bench.make_model(dev)
cfgd = dict(bench.CONFIGS["cfg1"]); cfgd["B"] = B; cfgd["T"] = T
inp = bench.make_inputs(cfgd, B)
kw = dict(fake_speaker=inp["fs"].to(dev), fake_content=inp["fc"].to(dev), cfg_strength=3.0)
g = {k: inp[k].to(dev) for k in ("mu", "mask", "c", "z")}
for _ in range(n):
    out = m(g["mu"], g["mask"], steps, 1.0, g["c"], "euler", kw, z=g["z"])
torch.cuda.synchronize()
print("done", float(out.abs().mean()), "launches", m.estimator.launch_count())
