"""Kernel-level timing of the conv-GEMM engine at the estimator's shapes (cfg1: BB=64, T=1000).
Usage: [STABLETTS_B200_TC2=0|1|force] python profiles/bench_gemm.py   (on a B200)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stabletts_b200 import CFMDecoder, _lib

dev = torch.device("cuda:0")
m = CFMDecoder(80, 80, 256, 80, 1024, 4, 6, 3, 0.1, 256).eval().to(dev)
m.estimator._prepare(torch.zeros(1, device=dev), 1, 8, 0)
lib, h = _lib.load_library(), m.estimator._handle
BB, T = int(os.environ.get("BG_BB", "64")), int(os.environ.get("BG_T", "1000"))
shapes = [("qkv", 256, 768, 1, 0), ("o_proj", 256, 256, 1, 1), ("conv_1", 256, 1024, 3, 0), ("conv_2", 1024, 256, 3, 1),
          ("lsc(K=512)", 512, 256, 3, 0), ("cond2", 1024, 1024, 3, 0)]
print(f"TC2={os.environ.get('STABLETTS_B200_TC2', 'default')} BB={BB} T={T}")
for name, cin, cout, k, epi in shapes:
    ms = C.c_float()
    rc = lib.st_bench_conv(h, BB, cin, cout, T, k, epi, 10, C.byref(ms))
    if rc:
        print(name, "FAILED", lib.st_last_error(h)); continue
    fl = 2.0 * BB * T * cin * cout * k
    print(f"{name:12s} K={cin*k:5d} N={cout:5d}  {ms.value*1e3:8.1f} us  {fl/ms.value/1e9:8.1f} TFLOP/s alg  ({3*fl/ms.value/1e9:7.1f} MMA)")
