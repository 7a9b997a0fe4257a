import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from stabletts_b200 import CFMDecoder, _lib
from oracle import estimator_ref as R
dev = torch.device("cuda:0")
m = CFMDecoder(80, 80, 256, 80, 1024, 4, 6, 3, 0.1, 256).eval().to(dev)
m.estimator._prepare(torch.zeros(1, device=dev), 1, 8, 0)
lib, h = _lib.load_library(), m.estimator._handle
s = torch.cuda.current_stream().cuda_stream
def run(mask, qkv):
    B, T = mask.shape
    out = torch.empty(B, T, 256, device=dev)
    _lib.check(lib, h, lib.st_test_attention(h, qkv.to(dev).data_ptr(), mask.to(dev).data_ptr(), out.data_ptr(), B, T, s), "att")
    torch.cuda.synchronize()
    return out.cpu()
def ref(mask, qkv):
    B, T = mask.shape
    q, k, v = [t.view(B, T, 4, 64).transpose(1, 2).double() for t in qkv.split(256, dim=-1)]
    q, k = R.rope_partial(q, 32), R.rope_partial(k, 32)
    am = mask[:, None, :, None] * mask[:, None, None, :]
    am = torch.zeros_like(am).masked_fill(am == 0, -torch.finfo(torch.float32).max).double()
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=am).transpose(1, 2).reshape(B, T, 256)
    return (o * mask[:, :, None]).float()
g = torch.Generator().manual_seed(3)
case = sys.argv[1]
if case == "holes":
    T = 200; lens = [200, 150]
    mask = (torch.arange(T)[None] < torch.tensor(lens)[:, None]).float()
    mask[0, 37:49] = 0; mask[0, 130] = 0; mask[1, 0:5] = 0
    qkv = torch.randn(2, T, 768, generator=g)
    o, r = run(mask, qkv), ref(mask, qkv)
    bad = torch.isnan(o).any(-1)
    print("nan rows b0:", bad[0].nonzero().flatten().tolist()[:40], "b1:", bad[1].nonzero().flatten().tolist()[:40])
    d = (o - r).abs().amax(-1)
    print("max err per batch", d.nan_to_num(9).amax(-1).tolist(), "worst rows", d.nan_to_num(9).argmax(-1).tolist())
elif case == "big":
    B, T = int(sys.argv[2]), int(sys.argv[3])
    mask = torch.ones(B, T)
    qkv = torch.randn(B, T, 768, generator=g)
    o = run(mask, qkv)
    print("big ok", float(o.abs().mean()), bool(torch.isnan(o).any()))
