"""Debug helper: the 20 x 1024-frame estimator call of smoke() with STABLETTS_B200_FUSE_LN=0/1 (separate processes), error of
each against the oracle on rows {0, 7, 19} and the difference between the two, located by frame / channel."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

if len(sys.argv) > 1:
    from oracle import weights
    from stabletts_b200 import CFMDecoder
    dev = torch.device("cuda:0")
    st = weights.make_state(0, 80)
    m = CFMDecoder(80, 80, 256, 80, 1024, 4, 6, 3, 0.1, 256).eval()
    m.estimator.load_state_dict(st, strict=True)
    m = m.to(dev)
    lens = [1024] * 20
    lens[7], lens[19] = 700, 1001
    big = weights.make_inputs(4, lens, 1024, 80)
    out = m.estimator(big["t"].to(dev), big["x"].to(dev), big["mask"].to(dev), big["mu"].to(dev), big["c"].to(dev)).cpu()
    torch.save(out, sys.argv[1])
    sys.exit(0)

outs = {}
for v in ("0", "1"):
    path = f"/tmp/fuse_{v}.pt"
    env = dict(os.environ, STABLETTS_B200_FUSE_LN=v)
    subprocess.run([sys.executable, __file__, path], env=env, check=True)
    outs[v] = torch.load(path)
from oracle import estimator_ref as R, weights
st = weights.make_state(0, 80)
lens = [1024] * 20
lens[7], lens[19] = 700, 1001
big = weights.make_inputs(4, lens, 1024, 80)
rows = [0, 7, 19]
with torch.inference_mode():
    ref = R.estimator_forward(st, big["t"], big["x"][rows], big["mask"][rows], big["mu"][rows], big["c"][rows])
for v in ("0", "1"):
    d = (outs[v][rows] - ref).abs()
    print("fuse", v, "max-rel vs oracle", float(d.max() / ref.abs().max()), "l2-rel", float((outs[v][rows] - ref).norm() / ref.norm()))
    idx = torch.nonzero(d > 0.2 * d.max())
    print("   worst elements (row, mel, frame):", idx[:12].tolist())
dd = (outs["1"] - outs["0"]).abs()
print("fuse1 vs fuse0: max abs", float(dd.max()), "rel", float(dd.max() / outs["0"].abs().max()))
fr = torch.nonzero(dd.amax(dim=1) > 0.2 * dd.max())
print("   frames with large difference (batch, frame):", fr[:40].tolist(), "count", len(fr))
