"""clock64 trace of one epilogue warp of the 2-CTA conv-GEMM (debug; run with STABLETTS_B200_EPI_TRACE=1):
    STABLETTS_B200_EPI_TRACE=1 python profiles/trace_epilogue.py
prints, per 32-channel chunk of the first tiles of CTA 0: wait for the accumulator, math, wait for the staging, stage+issue."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stabletts_b200 import _lib

dev = torch.device("cuda:0")
m = bench.make_model(dev)
m.estimator._prepare(torch.zeros(1, device=dev), 1, 8, 0)
lib, h = _lib.load_library(), m.estimator._handle
for name, (B, Cin, Cout, T, k, epi) in {"conv_1 (SiLU, split out)": (64, 256, 1024, 1000, 3, 2), "conv_2 (resid, f32+split)": (64, 1024, 256, 1000, 3, 1),
                                         "K=256 N=768 bias, split out": (64, 256, 768, 1000, 1, 0),
                                         "O (K=256 N=256, resid, f32 out, fused LN -> U)": (64, 256, 256, 1000, 1, 3),
                                         "O unfused (K=256 N=256, resid, f32+split out)": (64, 256, 256, 1000, 1, 1)}.items():
    ms = C.c_float()
    _lib.check(lib, h, lib.st_bench_conv(h, B, Cin, Cout, T, k, epi, 5, C.byref(ms)), "st_bench_conv")
    buf = (C.c_longlong * (4 * 16 * 8))()
    rc = lib.st_test_gemm_trace(buf)
    print(f"== {name}: {ms.value*1e3:.1f} us per launch (trace rc {rc})")
    prev_end = None
    for t in range(4):
        b = buf[(t * 16) * 8:(t * 16) * 8 + 8]
        if b[7]:
            print(f"  tile {t}: accumulator wait {b[6]-b[5]:7d} | whole tile (both passes) {b[7]-b[6]:7d}")
        for kc in range(16):
            d = buf[(t * 16 + kc) * 8:(t * 16 + kc) * 8 + 5]
            if d[4] == 0:
                continue
            gap = (d[0] - prev_end) if prev_end else 0
            post = f" | LN stats + tmem st {d[3]-d[4]:6d}" if d[3] else ""
            print(f"  tile {t} {'pass2 ' if kc >= 8 else ''}chunk {kc % 8}: gap {gap:6d} | acc wait {d[1]-d[0]:6d} | math {d[2]-d[1]:6d} | stage+issue {d[4]-d[2]:6d}{post} | total {max(d[4], d[3])-d[0]:6d}")
            prev_end = max(d[4], d[3])
