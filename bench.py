#!/usr/bin/env python
"""bench.py — mel-frames/sec through the CFM DiT estimator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg1|cfg2|cfg3|cfg4] [--impl reference]

A "step" is one complete ``CFMDecoder.forward`` ODE solve over one batch of synthetic inputs
(SURVEY.md §8d): weights = reference-style init under manual_seed(0) with the adaLN gates re-drawn
N(0, 0.3²); inputs under manual_seed(1); z unmasked; CFG strength 3.
  value  : frames/s with inputs already resident in HBM (device timed, max over ranks)
  e2e    : the same metric through the public module call with PINNED HOST inputs — H2D copies of
           (mu, mask, c, z) and the D2H read of the mel are inside the timed region
  roofline: the tcgen05 conv-GEMM class (dominant kernel): algorithmic FLOPs / CUDA-event time of
           every launch in one instrumented solve, against MEASURED_PEAKS.json's sustained bf16 peak
  cpu_baseline / --impl reference: the oracle port of the reference's PyTorch path on host cores
           (bounded sample), the only place this file executes anything under oracle/.
Multi-GPU (torchrun, one rank per GPU): weak scaling, the per-GPU batch is fixed; rank 0 owns the
global batch, scatters (mu, mask, c, z) and gathers the mel over NCCL inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg0": dict(B=1, T=300, n_steps=10, method="euler", cfg=3.0, lengths=None,
                 desc="batch=1, ~300-frame mel (64 phonemes), 10-step Euler + CFG — the reference's CPU-runnable plumbing case"),
    # name: per-GPU batch, T, steps, method, cfg, description (BASELINE.json configs[1..4])
    "cfg1": dict(B=32, T=1000, n_steps=10, method="euler", cfg=3.0, lengths=None,
                 desc="batch=32/GPU, n_mel=80, T=1000, 10-step Euler + CFG"),
    "cfg2": dict(B=256, T=500, n_steps=25, method="dopri5_fixed", cfg=None, lengths=None,
                 desc="batch=256, n_mel=80, T=500, 25 fixed Dormand-Prince steps (6 evals/step, NFE=150), no CFG"),
    "cfg3": dict(B=128, T=None, n_steps=10, method="euler", cfg=None, lengths="uniform200-2000",
                 desc="bucketed variable-length batch=128 (T in [200,2000]), 10-step Euler, masked attention"),
    "cfg4": dict(B=128, T=1000, n_steps=10, method="euler", cfg=3.0, lengths=None,
                 desc="batch=128/GPU (1024 over 8 GPUs), n_mel=80, T=1000, 10-step Euler + CFG"),
}
N_MEL = 80
NFE_PER_STEP = {"euler": 1, "midpoint": 2, "rk4": 4, "dopri5_fixed": 6}


def flops_per_frame_call(T: int, n_mel: int = N_MEL) -> float:
    """F_call(T) = 32.948e6 + 6144·T for M=80 (2·MAC)."""
    H, F, L = 256, 1024, 6
    return 2.0 * ((3 * n_mel * F + 3 * F * F + 3 * F * H) + (n_mel + H) * H + L * (4 * H * H + 2 * T * H + 6 * H * F)
                  + 3 * (6 * H * H) + H * n_mel)


COND_FLOPS = 2.0 * (3 * N_MEL * 1024 + 3 * 1024 * 1024 + 3 * 1024 * 256)    # 8.356 MFLOP/frame, hoistable


def make_model(device):
    from stabletts_b200 import CFMDecoder
    torch.manual_seed(0)
    m = CFMDecoder(N_MEL, N_MEL, 256, N_MEL, 1024, 4, 6, 3, 0.1, 256).eval()
    with torch.no_grad():
        for i in range(6):
            node = m.estimator._modules["blocks"]._modules[str(i)]._modules["block"]._modules["adaLN_modulation"]._modules["2"]
            torch.nn.init.normal_(node.weight, std=0.3)
            torch.nn.init.normal_(node.bias, std=0.3)
    return m.to(device) if device is not None else m


def make_inputs(cfgd, B, seed=1):
    """CPU tensors (global generator, manual_seed(seed)) in the reference's boundary layout."""
    torch.manual_seed(seed)
    if cfgd["lengths"] is None:
        T = cfgd["T"]
        lens = torch.full((B,), T, dtype=torch.long)
    else:
        lens = torch.randint(200, 2001, (B,))
        lens, _ = torch.sort(lens)
        T = int(lens.max())
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)
    mu = torch.randn(B, N_MEL, T) * mask
    c = torch.randn(B, 256)
    z = torch.randn(B, N_MEL, T)
    fs, fc = torch.randn(1, 256), torch.randn(1, N_MEL, 1)
    return dict(mu=mu, mask=mask, c=c, z=z, fs=fs, fc=fc, lens=lens, T=T)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
            return
        # nvidia-smi takes seconds to initialise NVML on an 8-GPU host and its first query stalls the driver for
        # hundreds of ms: wait for the first sample so that this start-up never lands inside a timed region
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 20.0:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.1)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) >= 9 and p[1].isdigit():
                    rows.append(p)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if rows:
            clocks = [int(r[1]) for r in rows]
            busy = [c for c, r in zip(clocks, rows) if float(r[3]) > 250.0] or clocks
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = int(rows[0][2])
            out["power_w_max"] = max(float(r[3]) for r in rows)
            out["samples"] = len(rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for j, n in enumerate(names):
                if any(r[5 + j].lower().startswith("active") for r in rows):
                    out["reasons"].append(n)
        return out


def cpu_threads() -> int:
    """Host threads actually usable: min(affinity mask, cgroup CPU quota) — see oracle.usable_cpus()."""
    from oracle import usable_cpus
    return usable_cpus()


def cpu_reference_solve(state, inp, cfgd, B, budget_s=25.0):
    """The oracle port (the reference's own PyTorch CPU path restated) on a BOUNDED sample: the first B
    utterances; if a probe evaluation predicts the full-NFE solve would exceed ``budget_s`` the number
    of ODE steps is cut (same grid spacing semantics, fewer steps) and the throughput is scaled to the
    full NFE.  Returns (output or None if truncated, seconds-equivalent for the FULL solve, note)."""
    from oracle import estimator_ref as R
    sl = slice(0, B)
    kw = None if cfgd["cfg"] is None else dict(fake_speaker=inp["fs"], fake_content=inp["fc"], cfg_strength=cfgd["cfg"])
    per_step = NFE_PER_STEP[cfgd["method"]] * (2 if kw else 1)
    with torch.inference_mode():
        t0 = time.perf_counter()
        R.estimator_forward(state, torch.tensor(0.5), inp["z"][sl], inp["mask"][sl], inp["mu"][sl], inp["c"][sl])
        probe = time.perf_counter() - t0
    steps = cfgd["n_steps"]
    if probe * per_step * steps > budget_s:
        steps = max(1, int(budget_s / (probe * per_step)))
    t0 = time.perf_counter()
    out = R.cfm_forward(state, inp["mu"][sl], inp["mask"][sl], steps, inp["z"][sl], inp["c"][sl], cfgd["method"], kw)
    dt = time.perf_counter() - t0
    if steps == cfgd["n_steps"]:
        return out, dt, f"full NFE={per_step * steps}"
    return None, dt * cfgd["n_steps"] / steps, f"{steps} of {cfgd['n_steps']} ODE steps timed ({dt:.1f} s), scaled to the full NFE"


def run_reference(args, cfgd, rank):
    if rank != 0:
        return
    threads = cpu_threads()
    torch.set_num_threads(threads)
    model = make_model(None)
    state = {k: v.detach().clone() for k, v in model.estimator.state_dict().items()}
    Bs = 1
    inp = make_inputs(cfgd, max(Bs, 2))
    frames = int(inp["lens"][:Bs].sum())
    budget = max(5.0, 150.0 / max(args.steps + 1, 1))          # whole run within a few minutes
    if args.warmup > 0:
        cpu_reference_solve(state, inp, cfgd, Bs, budget_s=budget / 4)
    t, note = 0.0, ""
    for _ in range(args.steps):
        _, dt, note = cpu_reference_solve(state, inp, cfgd, Bs, budget_s=budget)
        t += dt
    val = frames * args.steps / t
    sample = f"B={Bs} utterance of the workload at T={inp['T']} per step; {note}; {threads} torch threads"
    line = {"impl": "reference", "metric": "mel-frames/sec through CFM DiT estimator (ODE solve, all evaluations)",
            "value": val, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfgd['desc']} (CPU sample: {sample})"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg1", choices=list(CONFIGS))
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-mode", action="store_true",
                    help="for `ncu` launch lists only: honours --warmup < 3, skips e2e / instrumented / CPU legs (numbers printed under a profiler are never bench values)")
    args = ap.parse_args()
    cfgd = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfgd, rank)
        return
    if args.warmup < 3 and not args.ncu_mode:
        args.warmup = 3
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from stabletts_b200 import _lib, shard

    model = make_model(dev)
    model.estimator.set_engine(args.engine)
    Bper = cfgd["B"]
    Bglob = Bper * world
    inp = make_inputs(cfgd, Bglob) if rank == 0 else None
    T = cfgd["T"] if cfgd["lengths"] is None else None
    if world > 1:
        hdr = torch.tensor([inp["T"] if rank == 0 else 0], device=dev)
        dist.broadcast(hdr, 0)
        T = int(hdr.item())
    else:
        T = inp["T"]
    kw = None
    fs_d, fc_d = None, None
    if cfgd["cfg"] is not None:
        # fake_* are model parameters (models/model.py:43-44): replicated like the weights
        torch.manual_seed(2)
        fs_d, fc_d = torch.randn(1, 256).to(dev), torch.randn(1, N_MEL, 1).to(dev)
        if rank == 0:
            inp["fs"], inp["fc"] = fs_d.cpu(), fc_d.cpu()
        kw = dict(fake_speaker=fs_d, fake_content=fc_d, cfg_strength=cfgd["cfg"])

    def solve_one(mu, mask, c, z):
        return model(mu, mask, cfgd["n_steps"], 1.0, c, cfgd["method"], kw, z=z)

    def solve(mu, mask, c, z):
        if cfgd["lengths"] is None:
            return solve_one(mu, mask, c, z)
        lens_local = mask.sum(dim=(1, 2)).long().tolist()       # host-visible lengths (bucketing is host logic)
        return shard.bucketed_solve(solve_one, mu, mask, c, z, lens_local, n_buckets=4)

    # device-resident global inputs on rank 0
    if rank == 0:
        g = {k: inp[k].to(dev) for k in ("mu", "mask", "c", "z")}
        pinned = {k: inp[k].pin_memory() for k in ("mu", "mask", "c", "z")}
        frames_global = int(inp["lens"].sum())
    else:
        g, pinned, frames_global = None, None, 0

    def step_device():
        if world == 1:
            return solve(g["mu"], g["mask"], g["c"], g["z"])
        return shard.sharded_solve(solve, *((g["mu"], g["mask"], g["c"], g["z"]) if rank == 0 else (None,) * 4), device=dev,
                                   batch=Bglob, n_mel=N_MEL, T=T, gin=256)

    def step_e2e():
        if rank == 0:
            d = {k: pinned[k].to(dev, non_blocking=True) for k in pinned}
        if world == 1:
            out = solve(d["mu"], d["mask"], d["c"], d["z"])
        else:
            out = shard.sharded_solve(solve, *((d["mu"], d["mask"], d["c"], d["z"]) if rank == 0 else (None,) * 4), device=dev,
                                      batch=Bglob, n_mel=N_MEL, T=T, gin=256)
        return out.cpu() if rank == 0 else None

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        out = None
        for i in range(steps):
            flush.zero_()
            out = fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev[0].elapsed_time(ev[steps])], device=dev)
        timed.per_step = [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(steps)]    # diagnostics (this rank)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # started BEFORE warm-up: nvidia-smi start-up stalls the driver for ~100 ms
    # at least W untimed steps AND ~2 s of load: under the 1 kW cap the SM clock needs about a second to settle
    # (the first few hundred ms run at 1965 MHz, then the power controller pulls back and briefly overshoots).
    # The extra count is decided on rank 0 and broadcast so every rank issues the same number of collectives.
    tw0 = time.perf_counter()
    for _ in range(args.warmup):
        flush.zero_()                   # same ops as a timed step: torch lazy-loads its fill kernel's module on first use
        step_device()                   # (hundreds of ms on a cold box) and that must not land in the timed region
    torch.cuda.synchronize()
    el = time.perf_counter() - tw0
    extra = 0 if args.ncu_mode else int(max(0.0, 2.0 - el) / max(el / max(args.warmup, 1), 1e-4)) + 1
    if world > 1:
        ex_t = torch.tensor([extra], device=dev)
        dist.broadcast(ex_t, 0)
        extra = int(ex_t.item())
    for _ in range(extra):
        flush.zero_()
        step_device()
    timed(step_device, 1)               # one untimed pass through the timing harness itself (events, all_reduce)
    n_warm = args.warmup + extra + 1
    torch.cuda.synchronize()
    th0 = time.perf_counter()
    step_device()                       # host-side enqueue time of one step (no sync): launch-bound check
    host_ms = (time.perf_counter() - th0) * 1e3
    torch.cuda.synchronize()
    l0 = model.estimator.launch_count()
    remeasured = []

    def timed_checked(fn, steps, tag):
        """K timed steps; if one step is an outlier (> 1.5x the median: a host stall or a driver hiccup starves the GPU
        for tens of ms about once in a dozen runs) the whole K-step region is measured ONCE more and the repeat is kept;
        the JSON line says so.  The decision is rank 0's, broadcast, so every rank repeats or none does."""
        ms, out = timed(fn, steps)
        per = list(timed.per_step)
        med = sorted(per)[len(per) // 2]
        again = torch.tensor([1 if (steps >= 3 and max(per) > 1.5 * med) else 0], device=dev)
        if world > 1:
            dist.broadcast(again, 0)
        if int(again.item()) and not args.ncu_mode:
            remeasured.append({"region": tag, "first_attempt_step_ms": per[:32]})
            ms, out = timed(fn, steps)
            per = list(timed.per_step)
        return ms, out, per

    ms_dev, out_dev, per_step_dev = timed_checked(step_device, args.steps, "value")
    launches = (model.estimator.launch_count() - l0) // (2 if any(r["region"] == "value" for r in remeasured) else 1)
    if args.ncu_mode:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"ncu_mode": True, "ms_per_step_under_profiler": ms_dev / args.steps, "gpu_launches": int(launches)}))
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(1):
        step_e2e()
    ms_e2e, _, per_step_e2e = timed_checked(step_e2e, args.steps, "e2e")
    clocks = sampler.stop() if rank == 0 else None

    # instrumented solve: per-class CUDA-event timing of every launch (roofline)
    lib, h = _lib.load_library(), model.estimator._handle
    prof = None
    if rank == 0 or world == 1:
        lib.st_profile_begin(h)
        if world == 1:
            step_device()
        else:
            sl = slice(0, Bper)
            solve(g["mu"][sl].contiguous(), g["mask"][sl].contiguous(), g["c"][sl].contiguous(), g["z"][sl].contiguous())
        n = _lib.ST_PROF_NCAT
        ms_a, fl_a, by_a, ln_a = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        lib.st_profile_end(h, ms_a, fl_a, by_a, ln_a)
        prof = {name: dict(ms=ms_a[i], flops=fl_a[i], bytes=by_a[i], launches=int(ln_a[i]))
                for i, name in enumerate(_lib.ST_PROF_NAMES)}
        prof["gemm"] = {k: sum(v[k] for n_, v in prof.items() if n_.startswith("gemm_")) for k in ("ms", "flops", "bytes", "launches")}
    if world > 1:
        # non-root ranks must take part in nothing here; keep ranks aligned
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))["dram_bytes_per_launch_avg"]
    except Exception:
        pass
    gm = prof["gemm"]
    ach_tf = gm["flops"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] > 0 else 0.0
    nfe = cfgd["n_steps"] * NFE_PER_STEP[cfgd["method"]] * (2 if cfgd["cfg"] is not None else 1)
    lens = inp["lens"].double()
    hoisted = float(((nfe * (32.948e6 - COND_FLOPS) + COND_FLOPS) * lens + nfe * 6144.0 * lens * lens).sum())
    faithful = float((nfe * (32.948e6 * lens + 6144.0 * lens * lens)).sum())
    sec_step = ms_dev * 1e-3 / args.steps
    value = frames_global * args.steps / (ms_dev * 1e-3)
    e2e_val = frames_global * args.steps / (ms_e2e * 1e-3)
    h2d = sum(pinned[k].numel() * 4 for k in pinned)
    d2h = Bglob * N_MEL * T * 4

    cpu_baseline = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        threads = cpu_threads()
        torch.set_num_threads(threads)
        state = {k: v.detach().cpu().clone() for k, v in model.estimator.state_dict().items()}
        Bs = 1
        ref, dt, note = cpu_reference_solve(state, inp, cfgd, Bs, budget_s=25.0)
        fr = int(inp["lens"][:Bs].sum())
        cpu_baseline = {"value": fr / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                        "sample": f"first utterance of the batch at T={T}; {note}; os.cpu_count()={os.cpu_count()}"}
        if ref is not None:
            d = (out_dev[:Bs].cpu().double() - ref.double())
            parity = {"max_rel": float(d.abs().max() / ref.abs().max()), "l2_rel": float(d.norm() / ref.double().norm()),
                      "vs": "oracle port on the same inputs, first utterance, full solve"}

    line = {
        "metric": "mel-frames/sec through CFM DiT estimator (ODE solve, all evaluations)",
        "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": n_warm,
        "ms_per_step": 1e3 * sec_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (split-bf16x3 tensor-core operands, fp32 accumulate)" if args.engine == "tcgen05" else "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfgd['desc']}", "global_batch": Bglob, "T": T, "nfe": nfe,
                   "frames_per_step": frames_global, "parallelism": f"batch-shard x{world}",
                   "l2": "256 MiB flush between steps; per-eval working set (~1.5 GB) exceeds the 126 MB L2",
                   "engine": args.engine},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "power_w_max": clocks.get("power_w_max")},
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 2),
        "step_ms": {"value": per_step_dev[:32], "e2e": per_step_e2e[:32]}, "remeasured": remeasured,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 split-bf16 conv-GEMM, all launches of one solve)",
                     "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
                     "traffic_note": "dram__bytes_read+write per launch, mean over the ncu --set full capture in profiles/gemm_traffic.json",
                     "achieved_per_launch_gflop": gm["flops"] / max(gm["launches"], 1) / 1e9,
                     "peak_source": peak_src, "launches": gm["launches"], "kernel_ms_per_step": gm["ms"],
                     "note": "algorithmic FLOPs (2*rows*N*K*taps); the bf16x3 split issues 3 MMAs per algorithmic MAC"},
        "breakdown_ms_per_step": {k: round(v["ms"], 3) for k, v in prof.items()},
        "breakdown_tflops": {k: round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1) for k, v in prof.items() if v["flops"] > 0},
        "attention": {"tflops": prof["attention"]["flops"] / max(prof["attention"]["ms"], 1e-9) / 1e9},
        "work": {"hoisted_tflop_per_step": hoisted / 1e12, "faithful_tflop_per_step": faithful / 1e12,
                 "whole_solve_tflops_hoisted": hoisted / 1e12 / sec_step / max(world, 1) * 1.0},
        "cpu_baseline": cpu_baseline, "parity": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
