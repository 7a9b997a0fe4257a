"""Debug probe for the split-K path (STABLETTS_B200_DEBUG=2 synchronises around every GEMM of a solve and prints its shape):
python tests/diagnostics/probe_splitk.py B T [cfg]   — one 2-step Euler solve of seeded inputs, prints DONE when the device finished."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from oracle import cases, weights      # test infrastructure: seeded weights / inputs only
from stabletts_b200 import CFMDecoder

B, T = int(sys.argv[1]), int(sys.argv[2])
cfg = len(sys.argv) > 3 and sys.argv[3] == "cfg"
dev = torch.device("cuda:0")
m = CFMDecoder(80, 80, 256, 80, 1024, 4, 6, 3, 0.1, 256).eval()
m.estimator.load_state_dict(weights.make_state(cases.WEIGHT_SEED, 80), strict=True)
m = m.to(dev)
lens = [T - 7 * i for i in range(B)]
inp = weights.make_inputs(11, lens, T)
kw = None
if cfg:
    fs, fc = weights.make_cfg_params(cases.CFG_SEED)
    kw = dict(fake_speaker=fs.to(dev), fake_content=fc.to(dev), cfg_strength=3.0)
out = m(inp["mu"].to(dev), inp["mask"].to(dev), 2, 1.0, inp["c"].to(dev), "euler", kw, z=inp["x"].to(dev))
torch.cuda.synchronize()
print("DONE", B, T, cfg, float(out.abs().mean()), flush=True)
