#!/usr/bin/env python
"""bench.py — headline benchmark of the quantized Linear decode path on B200.

Metric (BASELINE.json): FP8/FP4 Linear decode tokens/s + achieved HBM GB/s.

Workload at N=1 (BASELINE.json configs[1]): PerChannelFp8 decode GEMV at Llama-3.1-8B layer shapes —
one "step" is one decode pass of M tokens (default M=1) through the Linear stack of the model's 32
MLP blocks: gate 4096->14336, up 4096->14336, down 14336->4096 per layer = 96 launches over 96
DISTINCT weight matrices (5.64 GB streamed per step, 45x the 126 MB L2, so every step reads HBM).
At N>1 the same stack is tensor-parallel (gate/up column-parallel, down row-parallel + NCCL
all-reduce): total work fixed -> "scaling": "strong".

  value      tokens/s with inputs resident in HBM (CUDA-graph replay of the launcher calls)
  e2e        same metric through LinearStack.forward_host: pinned H2D of the step's activations,
             the 96 launches, D2H of the result inside the timed region
  roofline   algorithmic bytes / measured duration of the GEMV kernel vs the measured HBM peak
  cpu_baseline / --impl reference
             the reference's CPU Linear (CpuLinearOp::forwardNaive restated in oracle/, FP32,
             long double accumulate, 1 thread — the reference does not thread batch <= 100)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HIDDEN, FFN, LAYERS = 4096, 14336, 32          # Llama-3.1-8B (Llama.Presets.ixx:91-96)
WORKLOADS = {
    # name: (hidden, ffn, layers, policy-name)
    "llama3.1-8b-mlp-fp8": (4096, 14336, 32, "fp8"),
    "gemma4-12b-mlp-fp4": (3840, 15360, 48, "fp4"),
    "llama3-70b-mlp-fp4": (8192, 28672, 16, "fp4"),
}
# --tokens > 16 turns any workload into the batched (prefill) measurement of BASELINE.json configs[3]:
# same Linear stack, a 2048-token batch per step, roofline bound = tensor (useful flops 2*M*N*K).


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (NVML, 50 ms period)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event(); self._t = None; self._nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()
        return self

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def stop(self) -> dict:
        self._stop.set()
        if self._t: self._t.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# CPU baseline (oracle port of CpuLinearOp::forwardNaive) — the only place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
def cpu_reference_leg(hidden: int, ffn: int, layers: int, M: int, steps: int, warmup: int,
                      budget_s: float = 20.0) -> dict:
    import numpy as np
    from oracle import oracle as O
    rng = np.random.default_rng(1234)
    # one layer's three FP32 matrices (the reference CPU Linear is FP32/unquantized: CpuLinearOp.ixx)
    Wg = (rng.standard_normal((ffn, hidden), dtype=np.float32) / np.float32(hidden ** 0.5))
    Wu = (rng.standard_normal((ffn, hidden), dtype=np.float32) / np.float32(hidden ** 0.5))
    Wd = (rng.standard_normal((hidden, ffn), dtype=np.float32) / np.float32(ffn ** 0.5))
    x = np.random.default_rng(99).standard_normal((M, hidden), dtype=np.float32)

    def one_layer(h):
        g = O.cpu_linear_forward(h, Wg, None, "auto")
        O.cpu_linear_forward(h, Wu, None, "auto")
        return O.cpu_linear_forward(g, Wd, None, "auto")

    # a "step" of the sample = ONE layer triple (1/layers of the real step), scaled afterwards
    t0 = time.perf_counter(); one_layer(x); first = time.perf_counter() - t0
    max_steps = max(1, int(budget_s / max(first, 1e-6)))
    w = min(warmup, max(0, max_steps // 8)); k = max(1, min(steps, max_steps - w))
    for _ in range(w): one_layer(x)
    times = []
    for _ in range(k):
        t0 = time.perf_counter(); one_layer(x); times.append(time.perf_counter() - t0)
    per_layer = sorted(times)[len(times) // 2]
    step_s = per_layer * layers
    val = M / step_s
    return {"value": val, "ms_per_step": step_s * 1e3, "steps": k, "warmup": w,
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": 1, "kind": "port",
                             "host_cores": os.cpu_count(),
                             "sample": f"{k} timed passes of ONE layer triple (gate,up,down = "
                                       f"{(2 * ffn * hidden + hidden * ffn) / 1e6:.0f} M MAC, FP32, long-double "
                                       f"accumulate, 1 thread as in CpuLinearOp::forwardNaive) x {layers} layers "
                                       f"extrapolated; median {per_layer * 1e3:.1f} ms/layer"}}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hidden, ffn, layers, pol = WORKLOADS[args.workload]
    r = cpu_reference_leg(hidden, ffn, layers, min(args.tokens, 16), args.steps, args.warmup, budget_s=60.0)
    line = {
        "impl": "reference", "metric": "linear_decode_tokens_per_s", "value": r["value"], "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, hidden, ffn, layers, args.tokens),
                   "note": "reference CPU Linear is FP32/unquantized (CpuLinearOp.ixx); same shapes, same M"},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(key, hidden, ffn, layers, M) -> str:
    return (f"{key}: {layers} layers x (gate {hidden}->{ffn}, up {hidden}->{ffn}, down {ffn}->{hidden}), "
            f"{'batched/prefill' if M > 16 else 'decode'} M={M}, all {3 * layers} weight matrices distinct")


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from mila_b200 import _lib
    from mila_b200.linear import PerChannelFp8, PerGroupFp4
    from mila_b200.stack import LinearStack

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "NONE"          # VERSION and WARN print "NCCL version ..." on stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()      # fail loudly if the extension is missing

    hidden, ffn, layers, pol = WORKLOADS[args.workload]
    policy = PerChannelFp8() if pol == "fp8" else PerGroupFp4(128)
    M = args.tokens
    mode = args.mode if M <= 16 else "launches"
    stack = LinearStack(hidden, ffn, layers, policy, M, dev, rank=rank, world=world,
                        group=dist.group.WORLD if world > 1 else None, allreduce=args.allreduce,
                        mode=mode, fuse_gate_up=args.fuse_gate_up)
    gen = torch.Generator(device="cpu"); gen.manual_seed(99)
    stack.x_host.copy_(torch.randn((M, hidden), generator=gen).to(torch.bfloat16))
    stack.set_input(stack.x_host.to(dev))
    if M > 16:
        _lib.check(_lib.lib().milab200_reserve_prefill(M, max(hidden, ffn)), "reserve_prefill")
    stack.capture()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k, w):
        for _ in range(w): fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k): fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms / k

    sampler = ClockSampler(local).start() if rank == 0 else None
    _lib.reset_launch_count()
    ms_dev = timed(stack.step, args.steps, max(args.warmup, 3))
    launches = stack.launches_per_step * args.steps
    ms_e2e = timed(lambda: (stack.forward_host(), torch.cuda.current_stream().synchronize()),
                   args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if sampler else None

    # sanity: the result of the last step is finite and non-trivial (guards "timed nothing")
    out = stack.y_host.float()
    assert torch.isfinite(out).all() and float(out.abs().max()) > 0

    if world > 1:
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if rank != 0:
        _finish(world)
        return

    alg_bytes = stack.algorithmic_bytes_per_step()          # per rank
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    peaks = json.loads(peaks_path.read_text()) if peaks_path.exists() else {}
    prefill = M > 16
    if prefill:
        # compute-bound regime: useful flops of this rank / time vs the measured cuBLAS BF16 dense rate
        # (sustained figure: the kernels are timed inside a long step).  The two-plane E4M3 MMA does 2x
        # the FP8-rate work per useful flop, so BF16 dense is the matching denominator.
        flops = sum(2.0 * M * qw.N * qw.K for t in stack.w for qw in t)
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0)); peak_src = "measured" if peaks else "fallback"
        achieved = flops / (ms_dev * 1e-3) / 1e12
    else:
        peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured" if peaks else "fallback"
        achieved = alg_bytes / (ms_dev * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try: traffic = (json.loads(tp.read_text()).get(f"{args.workload}:M{M}") or {}).get("traffic")
        except Exception: traffic = None

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_leg(hidden, ffn, layers, M, 20, 1, budget_s=15.0)["cpu_baseline"]

    line = {
        "metric": "linear_prefill_tokens_per_s" if prefill else "linear_decode_tokens_per_s",
        "value": M / (ms_dev * 1e-3), "unit": "tokens/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": _dtype_of(_lib.last_kernel(), pol),
        "data": "synthetic (random-init randn/sqrt(K) weights quantized on device, randn activations)",
        "config": {"workload": workload_name(args.workload, hidden, ffn, layers, M),
                   "parallelism": (f"tp{world} (gate/up column-parallel, down row-parallel, all-reduce: "
                                   f"{'nccl' if (args.allreduce == 'nccl' or M > 16) else 'fused in the GEMV epilogue over NVLink peer memory'})")
                                  if world > 1 else "single",
                   "l2": f"inputs larger than L2: {stack.weight_bytes() / 1e9:.2f} GB of weights streamed per step per GPU",
                   "timing": "CUDA events around graph replays, max over ranks"},
        "gpu_launches": int(launches),
        "launches_per_step": int(stack.launches_per_step),
        "e2e": {"value": M / (ms_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": M * hidden * 2, "d2h_bytes_per_step": M * hidden * 2},
        "roofline": ({"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                      "frac": achieved / peak, "peak_source": peak_src + " cuBLAS bf16 dense, sustained", "traffic": traffic,
                      "kernel": _lib.last_kernel(), "frac_of_nominal_2250_bf16": achieved / 2250.0}
                     if prefill else
                     {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                      "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                      "kernel": _lib.last_kernel(), "algorithmic_bytes_per_step": alg_bytes,
                      "frac_of_nominal_8TBs": achieved / 8000.0}),
        "clocks": clocks,
    }
    if cpu: line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    _finish(world)


def _dtype_of(kernel: str, pol: str) -> str:
    """The arithmetic the dominant kernel computes in (not a precision claim: results are FP32-exact weights
    times exactly-split BF16 activations, FP32 accumulate, one BF16 rounding)."""
    if kernel.startswith("decode_mx4"):
        return "e2m1 packed weights x eight exact e2m1 digit planes of the activations, tcgen05 kind::mxf4 (unit scale factors), f32 accumulate"
    w = "e4m3" if pol == "fp8" else "e2m1"
    return f"{w} weights x two exact e4m3 activation planes, tcgen05 kind::f8f6f4, f32 accumulate"


def _finish(world: int) -> None:
    """Leave without tearing NCCL down: destroying a process group whose collectives are held by a
    captured CUDA graph was seen to hang on exit (round 1, torchrun N=2), after the line was printed."""
    sys.stdout.flush(); sys.stderr.flush()
    if world > 1:
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="llama3.1-8b-mlp-fp8", choices=list(WORKLOADS))
    ap.add_argument("--tokens", type=int, default=1, help="tokens per step: 1..16 decode, > 16 batched/prefill (e.g. 2048)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="chain", choices=["chain", "launches"],
                    help="decode (M <= 16): the stack as ONE persistent chained launch (milab200_chain_*) or one launch per Linear")
    ap.add_argument("--fuse-gate-up", action="store_true",
                    help="gate and up as ONE Linear with the GLU in its epilogue (Mila's fc_gate_up + SwiGLU dataflow)")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1 decode: all-reduce fused into the row-parallel GEMV epilogue (NVLink peer memory) or NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
