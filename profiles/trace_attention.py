"""Per-key-block clock64 trace of one attention CTA (STABLETTS_B200_ATT_TRACE=1): where a block's time goes.
Not a benchmark.  Usage: STABLETTS_B200_ATT_TRACE=1 python profiles/trace_attention.py [B] [T]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stabletts_b200 import CFMDecoder, _lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
m = CFMDecoder(80, 80, 256, 80, 1024, 4, 6, 3, 0.1, 256).eval().to(dev)
m.estimator._prepare(torch.zeros(1, device=dev), B, T, 0)
lib, h = _lib.load_library(), m.estimator._handle
s = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(3)
qkv = torch.randn(B, T, 768, generator=g).to(dev)
mask = torch.ones(B, T, device=dev)
out = torch.empty(B, T, 256, device=dev)
for _ in range(3):
    _lib.check(lib, h, lib.st_test_attention(h, qkv.data_ptr(), mask.data_ptr(), out.data_ptr(), B, T, s), "att")
torch.cuda.synchronize()
buf = (C.c_longlong * (32 * 16))()
assert lib.st_test_attention_trace(buf) == 0
tr = [[buf[j * 16 + k] for k in range(16)] for j in range(32)]
t0 = min(x for row in tr for x in row if x)
names = ["sm:wait_S", "sm:S_ready", "sm:max_done", "sm:xchg_done", "sm:exp_done", "sm:P_stored", "sm:arrived", "-",
         "mma:loop_top", "mma:P_ready", "mma:PV_issued", "mma:K_ready", "mma:S_issued"]
print("block | " + " | ".join(names))
for j, row in enumerate(tr):
    if not any(row): break
    print(f"{j:5d} | " + " | ".join(f"{(x - t0) if x else -1:8d}" for x in row[:13]))
print("softmax per block: wait_S, ld+max, xchg, exp, st, arrive ; mma: wait_P, issue_PV, wait_K, issue_S")
for j, row in enumerate(tr):
    if not any(row) or j == 0: continue
    s_ = row
    print(f"{j:5d} | S-wait {s_[1]-s_[0]:6d} ld+max {s_[2]-s_[1]:5d} xchg {s_[3]-s_[2]:5d} exp {s_[4]-s_[3]:5d} st {s_[5]-s_[4]:5d} arr {s_[6]-s_[5]:5d} "
          f"| period {s_[6]-tr[j-1][6]:6d} || P-wait {s_[9]-s_[8]:6d} PV {s_[10]-s_[9]:5d} K-wait {max(0,s_[11]-s_[10]):5d} S {max(0,s_[12]-s_[11]):5d}")
