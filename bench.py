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
# CPU legs (oracle port of CpuLinearOp::forwardNaive) — the only place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
def cpu_stack_leg(hidden: int, ffn: int, layers: int, M: int, steps: int, warmup: int, budget_s: float) -> dict:
    """The reference's CPU Linear (CpuLinearOp.ixx:384-411 restated in oracle/: FP32, `long double` accumulate, ONE
    thread — the reference does not thread batch <= 100 and builds with OpenMP off, CMakeLists.txt:78) over the SAME
    stack as the GPU arm: every timed step is one full pass of layers x (gate, up, down).  The three FP32 matrices of one
    layer are reused for every layer (96 distinct FP32 matrices would be 22.6 GB of host memory; at 235 MB each they
    never stay in cache, so reuse does not help the CPU).  Steps are reduced, and said so, if K+W full passes would not
    fit `budget_s`."""
    import numpy as np
    from oracle import oracle as O
    rng = np.random.default_rng(1234)
    Wg = (rng.standard_normal((ffn, hidden), dtype=np.float32) / np.float32(hidden ** 0.5))
    Wu = (rng.standard_normal((ffn, hidden), dtype=np.float32) / np.float32(hidden ** 0.5))
    Wd = (rng.standard_normal((hidden, ffn), dtype=np.float32) / np.float32(ffn ** 0.5))
    x = np.random.default_rng(99).standard_normal((M, hidden), dtype=np.float32)

    def one_step():
        h = x
        for _ in range(layers):
            g = O.cpu_linear_forward(h, Wg, None, "auto")
            O.cpu_linear_forward(h, Wu, None, "auto")
            h = O.cpu_linear_forward(g, Wd, None, "auto")
            h = h / np.float32(max(float(np.abs(h).max()), 1e-30))      # keep the chain finite (host glue, ~us)
        return h

    t0 = time.perf_counter(); one_step(); first = time.perf_counter() - t0
    fit = max(1, int(budget_s / max(first, 1e-6)) - 1)                   # the probe pass counts as a warm-up
    w = max(0, min(warmup, fit // 4) - 1)
    k = max(1, min(steps, fit - w))
    for _ in range(w): one_step()
    times = []
    for _ in range(k):
        t0 = time.perf_counter(); one_step(); times.append(time.perf_counter() - t0)
    step_s = sum(times) / len(times)
    val = M / step_s
    mac = layers * (2 * ffn * hidden + hidden * ffn) * M
    return {"value": val, "ms_per_step": step_s * 1e3, "steps": k, "warmup": w + 1,
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                             "sample": f"{k} timed full passes (+{w + 1} warm-up) of the whole {3 * layers}-Linear stack, "
                                       f"{mac / 1e9:.2f} G MAC per pass, FP32, long-double accumulate, 1 thread as in "
                                       f"CpuLinearOp::forwardNaive (one layer's matrices reused for all layers); "
                                       f"{step_s:.2f} s per pass"}}


def cpu_context_rows(budget_s: float = 6.0) -> dict:
    """BASELINE.json configs[0] as stated (one CPU Linear, FP32, 2048 -> 8192, batch 1, randn/sqrt(K) weights seed 1234,
    x seed 99: BASELINE.md §4) through the reference's path, plus the two labelled NON-reference context rows of
    SURVEY.md §8d: the same loop with a float accumulator, and torch's CPU matmul on all cores."""
    import numpy as np
    from oracle import oracle as O
    K, N = 2048, 8192
    W = (np.random.default_rng(1234).standard_normal((N, K), dtype=np.float32) / np.float32(K ** 0.5))
    x = np.random.default_rng(99).standard_normal((1, K), dtype=np.float32)

    def best_median(fn, n=5, warm=3):
        for _ in range(warm): fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        ts.sort()
        return ts[0], ts[len(ts) // 2]
    out = {"config": "BASELINE.json configs[0]: single CPU Linear forward, FP32, 2048->8192, batch 1", "weight_MB": N * K * 4 / 1e6}
    b, m = best_median(lambda: O.cpu_linear_forward(x, W, None, "auto"))
    out["reference_forwardNaive_1thread"] = {"best_ms": b * 1e3, "median_ms": m * 1e3, "tokens_per_s": 1.0 / m, "kind": "port"}
    b, m = best_median(lambda: O.cpu_linear_forward(x, W, None, "context_f32acc"))
    out["context_float_accumulator_1thread"] = {"best_ms": b * 1e3, "median_ms": m * 1e3, "tokens_per_s": 1.0 / m,
                                                "kind": "context (not the reference): gcc -O2, float accumulate"}
    try:
        import torch
        Wt, xt = torch.from_numpy(W), torch.from_numpy(x)
        b, m = best_median(lambda: torch.matmul(xt, Wt.t()))
        out["context_torch_cpu_all_cores"] = {"best_ms": b * 1e3, "median_ms": m * 1e3, "tokens_per_s": 1.0 / m,
                                              "threads": torch.get_num_threads(), "kind": "context (not the reference)"}
    except Exception as e:                                                # pragma: no cover
        out["context_torch_cpu_all_cores"] = {"error": str(e)}
    return out


def base_config(key, hidden, ffn, layers, M, world, allreduce, fuse: bool = False) -> dict:
    """The `config` object both arms print (identical keys and values for the same flags)."""
    return {"workload": workload_name(key, hidden, ffn, layers, M, fuse),
            "parallelism": (f"tp{world}: gate/up column-parallel, down row-parallel + all-reduce" if world > 1 else "single")}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hidden, ffn, layers, pol = WORKLOADS[args.workload]
    M = min(args.tokens, 16)
    r = cpu_stack_leg(hidden, ffn, layers, M, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "linear_decode_tokens_per_s", "value": r["value"], "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": base_config(args.workload, hidden, ffn, layers, args.tokens, int(os.environ.get("WORLD_SIZE", "1")), args.allreduce,
                              args.fuse_gate_up),
        "what": "the reference's own CPU Linear (FP32, unquantized: CpuLinearOp.ixx), same shapes and M, every step a full "
                "pass of the stack; steps/warmup are the passes actually timed",
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(key, hidden, ffn, layers, M, fuse: bool = False) -> str:
    if fuse:
        # Mila's own MLP dataflow: fc_gate_up is ONE Linear [2 ffn, hidden] followed by the GLU (Gemma.Block.ixx:347,
        # Llama.Block.ixx:883); here the GLU runs in that Linear's epilogue.  Same weight bytes as the three-Linear form.
        glu = "GeGLU" if key.startswith("gemma") else "SwiGLU"
        how = "as ONE launch" if M <= 16 else "as activation pre-pass + ONE GEMM"
        return (f"{key}: {layers} layers x MLP block (RMSNorm -> gate_up {hidden}->{2 * ffn} -> {glu} {how}, then down {ffn}->{hidden}), "
                f"{'batched/prefill' if M > 16 else 'decode'} M={M}, all {2 * layers} weight matrices distinct")
    return (f"{key}: {layers} layers x (gate {hidden}->{ffn}, up {hidden}->{ffn}, down {ffn}->{hidden}), "
            f"{'batched/prefill' if M > 16 else 'decode'} M={M}, all {3 * layers} weight matrices distinct")


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def pick_mode(requested: str, pol: str, M: int, hidden: int = 0, ffn: int = 0, world: int = 1, fuse: bool = False) -> str:
    """auto: the chained persistent launch where it measured faster — M <= 8 and Linears of >= ~40 MB per GPU (Llama-8B FP8
    58.7 MB: 1010 vs 880 tok/s; Llama-70B FP4 117 MB: 946 vs 841; profiles/r2j18_*, r2j20_*).  Per-Linear launches otherwise:
    a kernel boundary under programmatic dependent launch costs ~2.8 us, the chain's device-side dependency ~6 us, so
    small Linears (Gemma-12B FP4: 29.5 MB, 797 vs 827 tok/s) are better off as separate launches; and at M > 8 the
    activation pre-pass of the per-Linear route beats in-kernel conversion."""
    if M > 16 or fuse:
        return "launches"      # (fused gate|up: the RMSNorm prologue exists on the per-Linear routes only; the chain also loses there, r2k1/r2k2)
    if requested != "auto":
        return requested
    bytes_per_linear = hidden * ffn * (1.0 if pol == "fp8" else 0.53125) / max(world, 1)
    if world > 1:
        return "chain" if M <= 8 else "launches"          # tensor parallel: measured faster at every world size (r2j10)
    if pol == "fp4" and M == 1 and bytes_per_linear >= 25e6:
        return "chain"        # one-token epilogue of the packed-nibble scheme (r2k24): Gemma-12B FP4 942 chained vs 863 launched
    return "chain" if (M <= 8 and bytes_per_linear >= 40e6) else "launches"


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    return json.loads(p.read_text()) if p.exists() else {}


def traffic_of(key: str, M: int, mode: str = "launches"):
    """DRAM bytes (read + written) of ONE launch of the dominant kernel from an `ncu --set full` capture (profiles/traffic.json).
    A chained stack is one launch per step: its entry is the whole chain's traffic."""
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            t = json.loads(tp.read_text())
            return (t.get(f"{key}:M{M}:{mode}") or ({} if mode == "chain" else t.get(f"{key}:M{M}")) or {}).get("traffic")
        except Exception: return None
    return None


def roofline_record(stack, M: int, ms: float, kernel: str, key: str, mode: str = "launches") -> dict:
    pk = peaks()
    if M > 16:
        # compute-bound regime: useful flops / time vs the measured cuBLAS BF16 dense rate (sustained figure: the kernels
        # run inside a long step).  The exact two-plane E4M3 MMA does 2x the FP8-rate work per useful flop, so BF16
        # dense is the matching denominator; the nominal figures are given beside it.
        flops = sum(2.0 * M * qw.N * qw.K for t in stack.w for qw in t)
        peak = float(pk.get("bf16_tflops_sustained", 1400.0))
        ach = flops / (ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": ("measured" if pk else "fallback") + " cuBLAS bf16 dense, sustained", "traffic": traffic_of(key, M),
                "kernel": kernel, "frac_of_nominal_2250_bf16": ach / 2250.0, "frac_of_nominal_4500_fp8": ach / 4500.0}
    alg = stack.algorithmic_bytes_per_step()
    peak = float(pk.get("hbm_gbs", 6650.0))
    ach = alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "peak_source": "measured" if pk else "fallback", "traffic": traffic_of(key, M, mode), "kernel": kernel,
            "algorithmic_bytes_per_step": alg, "frac_of_nominal_8TBs": ach / 8000.0}


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

    def measure(key: str, M: int, steps: int, warmup: int, mode_req: str, e2e: bool, fuse: bool = False) -> dict:
        """Build the stack of workload `key`, capture it, time `steps` replays (device-resident inputs) and, if asked,
        `steps` end-to-end passes (pinned H2D + stack + D2H).  Clocks are sampled over both timed regions."""
        hidden, ffn, layers, pol = WORKLOADS[key]
        policy = PerChannelFp8() if pol == "fp8" else PerGroupFp4(128)
        mode = pick_mode(mode_req, pol, M, hidden, ffn, world, fuse)
        if world > 1 and args.allreduce == "nccl":
            mode = "launches"
        stack = LinearStack(hidden, ffn, layers, policy, M, dev, rank=rank, world=world,
                            group=dist.group.WORLD if world > 1 else None, allreduce=args.allreduce, mode=mode,
                            fuse_gate_up=fuse, glu_kind=(1 if key.startswith("gemma") else 2))
        gen = torch.Generator(device="cpu"); gen.manual_seed(99)
        stack.x_host.copy_(torch.randn((M, hidden), generator=gen).to(torch.bfloat16))
        stack.set_input(stack.x_host.to(dev))
        if M > 16:
            _lib.check(_lib.lib().milab200_reserve_prefill(M, max(hidden, ffn)), "reserve_prefill")
        stack.capture()
        sampler = ClockSampler(local).start() if rank == 0 else None
        ms_dev = timed(stack.step, steps, warmup)
        kernel = _lib.last_kernel()
        ms_e2e = None
        if e2e:
            ms_e2e = timed(lambda: (stack.forward_host(), torch.cuda.current_stream().synchronize()), steps, warmup)
            out = stack.y_host.float()
        else:
            out = stack.step().float().cpu()
        clocks = sampler.stop() if sampler else None
        # sanity: the result of the last step is finite and non-trivial (guards "timed nothing")
        assert torch.isfinite(out).all() and float(out.abs().max()) > 0
        rec = {"workload": key, "M": M, "mode": mode, "ms_per_step": ms_dev, "ms_e2e": ms_e2e, "kernel": kernel,
               "launches_per_step": int(stack.launches_per_step), "clocks": clocks, "hidden": hidden, "ffn": ffn,
               "layers": layers, "pol": pol, "weight_GB": stack.weight_bytes() / 1e9, "fuse": fuse,
               "roofline": roofline_record(stack, M, ms_dev, kernel, key, mode) if rank == 0 else None, "stack": stack}
        return rec

    steps, warmup = args.steps, max(args.warmup, 3)
    if args.norm_fast:
        _lib.set_option("rmsnorm_fast_reduction", 1)
    _lib.reset_launch_count()
    main = measure(args.workload, args.tokens, steps, warmup, args.mode, e2e=True, fuse=args.fuse_gate_up)
    stack = main["stack"]
    M, hidden, ffn, layers, pol = args.tokens, main["hidden"], main["ffn"], main["layers"], main["pol"]
    launches = stack.launches_per_step * steps * 2            # device-resident + end-to-end timed regions

    # N > 1: parity of the tensor-parallel arithmetic against the single-GPU result, outside the timed region
    tp_parity = None
    if world > 1 and M <= 16:
        from mila_b200.tp import tp_parity_record
        policy = PerChannelFp8() if pol == "fp8" else PerGroupFp4(128)
        tp_parity = tp_parity_record(stack.tp, policy, hidden, ffn, M)
        tp_parity["gate"] = 1e-2
        tp_parity["ok"] = bool(tp_parity["identical_bits_across_ranks"] and tp_parity["max_rel_err_rowabs"] <= 1e-2
                               and tp_parity["max_rel_err_rowabs_vs_fp32_dequant_gemm"] <= 1e-2)
    if world > 1:
        barrier()
    del stack; main["stack"] = None
    torch.cuda.empty_cache()

    # ---- extra records: the rest of BASELINE.json's configs, each with its own roofline and clocks ----
    extras = []
    if world > 1 and not args.no_extras and args.workload != "llama3-70b-mlp-fp4":
        # BASELINE.json configs[4]: tensor-parallel FP4 stack at Llama-3-70B shapes (8192 <-> 28672), M = 1, this world size
        try:
            r = measure("llama3-70b-mlp-fp4", 1, max(3, min(steps, 10)), 3, "auto", e2e=False)
            st70 = r.pop("stack", None)
            from mila_b200.tp import tp_parity_record
            par = tp_parity_record(st70.tp, PerGroupFp4(128), 8192, 28672, 1)
            del st70
            torch.cuda.empty_cache()
            extras.append({"name": f"llama3-70b-mlp-fp4:M1:tp{world}", "metric": "linear_decode_tokens_per_s",
                           "value": 1.0 / (r["ms_per_step"] * 1e-3), "unit": "tokens/s", "ms_per_step": r["ms_per_step"],
                           "mode": r["mode"], "launches_per_step": r["launches_per_step"], "weight_GB_per_gpu": r["weight_GB"],
                           "roofline": r["roofline"], "clocks": r["clocks"], "tp_parity": par})
        except Exception as e:
            extras.append({"name": f"llama3-70b-mlp-fp4:M1:tp{world}", "error": f"{type(e).__name__}: {e}"})
        barrier()
    if world == 1 and not args.no_extras:
        plan = [("llama3.1-8b-mlp-fp8", 16), ("gemma4-12b-mlp-fp4", 1), ("gemma4-12b-mlp-fp4", 16),
                ("llama3-70b-mlp-fp4", 1), ("llama3.1-8b-mlp-fp8", 2048), ("gemma4-12b-mlp-fp4", 2048)]
        for key, m in plan:
            if key == args.workload and m == args.tokens:
                continue
            try:
                r = measure(key, m, max(3, min(steps, 10)), 3, "auto", e2e=False)
                r.pop("stack", None)
                torch.cuda.empty_cache()
                extras.append({"name": f"{key}:M{m}", "metric": "linear_prefill_tokens_per_s" if m > 16 else "linear_decode_tokens_per_s",
                               "value": m / (r["ms_per_step"] * 1e-3), "unit": "tokens/s", "ms_per_step": r["ms_per_step"],
                               "mode": r["mode"], "launches_per_step": r["launches_per_step"], "weight_GB": r["weight_GB"],
                               "roofline": r["roofline"], "clocks": r["clocks"]})
            except Exception as e:                                        # an extra must never cost the headline line
                extras.append({"name": f"{key}:M{m}", "error": f"{type(e).__name__}: {e}"})
        # Mila's own MLP dataflow (ln_2 -> fc_gate_up -> GLU -> fc_down, Gemma.Block.ixx:209-210,347-349; Llama.Block.ixx:883):
        # RMSNorm + gate|up + GLU as ONE launch, then down — two launches per layer, same weight bytes, plus the norm.
        # RMS reduction: the default reference order (bit-identical to the kernel sequence) or the opt-in tree order.
        # At M = 2048 (FP8 weights) the front half is two launches: the activation pre-pass that normalises while it splits, and
        # the GEMM with the GLU in its epilogue (prefill_tc_kernel<fp8,cta_pair,glu>); the [M, 2 ffn] tensor is never written.
        for key, m, fast in [("gemma4-12b-mlp-fp4", 1, 0), ("gemma4-12b-mlp-fp4", 1, 1), ("llama3.1-8b-mlp-fp8", 1, 1),
                             ("llama3.1-8b-mlp-fp8", 16, 1), ("llama3.1-8b-mlp-fp8", 2048, 0)]:
            if True:
                name = f"{key}:M{m}:mlp_block" + (":norm_fast" if fast else "")
                try:
                    _lib.set_option("rmsnorm_fast_reduction", fast)
                    r = measure(key, m, max(3, min(steps, 10)), 3, "auto", e2e=False, fuse=True)
                    r.pop("stack", None)
                    torch.cuda.empty_cache()
                    hidden_, ffn_, layers_, _ = WORKLOADS[key]
                    extras.append({"name": name, "metric": "linear_prefill_tokens_per_s" if m > 16 else "linear_decode_tokens_per_s",
                                   "what": workload_name(key, hidden_, ffn_, layers_, m, True) +
                                           ("; RMS reduction in tree order (opt-in, rstd within FP32 ulps of the reference order)" if fast
                                            else "; RMS reduction in the reference's order: bit-identical to RMSNorm -> Linear -> GLU kernels"),
                                   "value": m / (r["ms_per_step"] * 1e-3), "unit": "tokens/s", "ms_per_step": r["ms_per_step"],
                                   "mode": r["mode"], "launches_per_step": r["launches_per_step"], "weight_GB": r["weight_GB"],
                                   "roofline": r["roofline"], "clocks": r["clocks"]})
                except Exception as e:
                    extras.append({"name": name, "error": f"{type(e).__name__}: {e}"})
                finally:
                    _lib.set_option("rmsnorm_fast_reduction", 1 if args.norm_fast else 0)
        # opt-in FP8-rate prefill mode: one per-token E4M3 activation plane (the reference's own W4A8 activation format),
        # reported separately against an FP8 dense peak measured on this box with cuBLASLt (torch._scaled_mm)
        fp8_peak = fp8_src = None
        try:
            _lib.set_option("prefill_act_planes", 1)
            r = measure("llama3.1-8b-mlp-fp8", 2048, max(3, min(steps, 10)), 3, "auto", e2e=False)
            r.pop("stack", None)
            torch.cuda.empty_cache()
            fp8_peak, fp8_src = measure_fp8_peak()
            ach = r["roofline"]["achieved"]
            extras.append({"name": "llama3.1-8b-mlp-fp8:M2048:act_planes=1", "metric": "linear_prefill_tokens_per_s",
                           "what": "opt-in lossy mode: ONE per-token E4M3 activation plane (CudaFp8Prefill.cu:116-165 format, gate "
                                   "1e-1 row_absmax) x raw E4M3 weights; the default two-plane mode above is the conforming path",
                           "value": 2048 / (r["ms_per_step"] * 1e-3), "unit": "tokens/s", "ms_per_step": r["ms_per_step"],
                           "roofline": {"bound": "tensor", "achieved": ach, "peak": fp8_peak, "unit": "TFLOP/s", "frac": ach / fp8_peak,
                                        "peak_source": fp8_src, "kernel": r["kernel"], "frac_of_nominal_4500_fp8": ach / 4500.0},
                           "clocks": r["clocks"]})
        except Exception as e:
            extras.append({"name": "llama3.1-8b-mlp-fp8:M2048:act_planes=1", "error": f"{type(e).__name__}: {e}"})
        finally:
            _lib.set_option("prefill_act_planes", 2)
        # the same one-plane format against FP4 weights: exactly what the reference's batched FP4 path does (W4A8,
        # LIN/CudaLinearOp.ixx:660-714), with the per-group FP32 promotion kept
        try:
            _lib.set_option("prefill_act_planes", 1)
            r = measure("gemma4-12b-mlp-fp4", 2048, max(3, min(steps, 10)), 3, "auto", e2e=False)
            r.pop("stack", None)
            torch.cuda.empty_cache()
            if fp8_peak is None:
                fp8_peak, fp8_src = measure_fp8_peak()
            ach = r["roofline"]["achieved"]
            extras.append({"name": "gemma4-12b-mlp-fp4:M2048:act_planes=1", "metric": "linear_prefill_tokens_per_s",
                           "what": "opt-in lossy mode: ONE per-token E4M3 activation plane x raw E2M1 weights (W4A8, the reference's own "
                                   "batched FP4 activation format, gate 1e-1 row_absmax); the summed-planes record above is the conforming path",
                           "value": 2048 / (r["ms_per_step"] * 1e-3), "unit": "tokens/s", "ms_per_step": r["ms_per_step"],
                           "roofline": {"bound": "tensor", "achieved": ach, "peak": fp8_peak, "unit": "TFLOP/s", "frac": ach / fp8_peak,
                                        "peak_source": fp8_src, "kernel": r["kernel"], "frac_of_nominal_4500_fp8": ach / 4500.0},
                           "clocks": r["clocks"]})
        except Exception as e:
            extras.append({"name": "gemma4-12b-mlp-fp4:M2048:act_planes=1", "error": f"{type(e).__name__}: {e}"})
        finally:
            _lib.set_option("prefill_act_planes", 2)
        try:
            extras.append(reference_gpu_record(args, timed))
        except Exception as e:
            extras.append({"name": "reference_gpu", "error": f"{type(e).__name__}: {e}"})

    if rank != 0:
        _finish(world)
        return

    cpu = ctx = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_stack_leg(hidden, ffn, layers, min(M, 16), 3, 1, budget_s=25.0)["cpu_baseline"]
        ctx = cpu_context_rows()

    prefill = M > 16
    cfg = base_config(args.workload, hidden, ffn, layers, M, world, args.allreduce, args.fuse_gate_up)
    line = {
        "metric": "linear_prefill_tokens_per_s" if prefill else "linear_decode_tokens_per_s",
        "value": M / (main["ms_per_step"] * 1e-3), "unit": "tokens/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": _dtype_of(main["kernel"], pol),
        "data": "synthetic (random-init randn/sqrt(K) weights quantized on device, randn activations)",
        "config": cfg,
        "method": {"mode": main["mode"] + (": the whole stack as ONE persistent chained launch (milab200_chain_*), replayed as a CUDA graph"
                                           if main["mode"] == "chain" else ": one launcher call per Linear, replayed as a CUDA graph"),
                   "all_reduce": (("nccl" if (args.allreduce == "nccl" or M > 16) else "fused in the GEMV epilogue over NVLink peer memory")
                                  if world > 1 else None),
                   "l2": f"inputs larger than L2: {main['weight_GB']:.2f} GB of weights streamed per step per GPU",
                   "timing": "CUDA events around graph replays, barrier + synchronize on both sides, max over ranks"},
        "gpu_launches": int(launches),
        "launches_per_step": int(main["launches_per_step"]),
        "e2e": {"value": M / (main["ms_e2e"] * 1e-3), "unit": "tokens/s", "ms_per_step": main["ms_e2e"],
                "h2d_bytes_per_step": M * hidden * 2, "d2h_bytes_per_step": M * hidden * 2},
        "roofline": main["roofline"],
        "clocks": main["clocks"],
    }
    if tp_parity is not None: line["tp_parity"] = tp_parity
    if cpu: line["cpu_baseline"] = cpu
    if ctx: line["cpu_context"] = ctx
    if extras: line["extra"] = extras
    print(json.dumps(line), flush=True)
    _finish(world)


def measure_fp8_peak():
    """Dense FP8 (E4M3 x E4M3 -> BF16) matmul rate of this box through cuBLASLt (torch._scaled_mm), 8192^3, sustained over
    ~1 s after warm-up — the denominator for the FP8-rate prefill mode.  Falls back to the nominal 4500 TFLOP/s."""
    import torch
    try:
        n = 8192
        a = torch.randn((n, n), device="cuda").to(torch.float8_e4m3fn)
        b = torch.randn((n, n), device="cuda").to(torch.float8_e4m3fn).t()
        one = torch.ones((), device="cuda")
        f = lambda: torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 40
        e0.record()
        for _ in range(k): f()
        e1.record(); torch.cuda.synchronize()
        return 2.0 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12, "measured here: torch._scaled_mm (cuBLASLt) e4m3 x e4m3 8192^3, 40 back to back"
    except Exception as e:                                                   # pragma: no cover
        return 4500.0, f"fallback: nominal dense FP8 ({type(e).__name__})"


def reference_gpu_record(args, timed) -> dict:
    """The reference's OWN decode kernels (cuda_matvec_decode_bf16_qfp8, CudaMatVecBias.Bf16.cu:527, compiled unmodified for
    sm_100a into oracle/_ref — BASELINE.md §4 'the number to beat on the same box') over the default workload's stack in
    the same harness: 96 launches over 96 distinct weight matrices captured in a CUDA graph, M = 1."""
    import ctypes
    import torch
    from mila_b200.linear import PerChannelFp8
    from mila_b200.stack import LinearStack
    from oracle import oracle as O                      # checker library timed as a baseline, never as the product
    if not O.ref_lib_path().exists():
        return {"name": "reference_gpu", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
    R = O.ref_lib()
    hidden, ffn, layers, _ = WORKLOADS["llama3.1-8b-mlp-fp8"]
    dev = torch.device("cuda", torch.cuda.current_device())
    stack = LinearStack(hidden, ffn, layers, PerChannelFp8(), 1, dev, mode="launches")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    x0 = torch.randn((1, hidden), device=dev).to(torch.bfloat16)
    stack.set_input(x0)

    def forward():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        cur = 0
        for (gate, up, down) in stack.w:
            hin, hout = stack.h[cur], stack.h[cur ^ 1]
            R.milaref_matvec_decode_bf16_qfp8(p(stack.g), p(hin), p(gate.weight), p(gate.scales), None, hidden, ffn, st)
            R.milaref_matvec_decode_bf16_qfp8(p(stack.u), p(hin), p(up.weight), p(up.scales), None, hidden, ffn, st)
            R.milaref_matvec_decode_bf16_qfp8(p(hout), p(stack.g), p(down.weight), p(down.scales), None, ffn, hidden, st)
            cur ^= 1
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        forward()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        forward()
    sampler = ClockSampler(torch.cuda.current_device()).start()
    ms = timed(g.replay, max(3, min(args.steps, 10)), 3)
    clocks = sampler.stop()
    alg = stack.algorithmic_bytes_per_step()
    peak = float(peaks().get("hbm_gbs", 6650.0))
    rec = {"name": "reference_gpu", "what": "Mila's own cuda_matvec_decode_bf16_qfp8 kernels recompiled unmodified for sm_100a, "
           "llama3.1-8b-mlp-fp8 stack, M=1, same graph harness", "value": 1.0 / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms,
           "launches_per_step": 3 * layers,
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / peak, "kernel": "reference matvec_decode_bf16_qfp8"},
           "clocks": clocks}
    del stack
    torch.cuda.empty_cache()
    return rec


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
    ap.add_argument("--fuse-gate-up", action="store_true",
                    help="gate and up as ONE Linear with the GLU in its epilogue (Mila's fc_gate_up dataflow): two Linears per layer")
    ap.add_argument("--norm-fast", action="store_true",
                    help="with --fuse-gate-up: RMSNorm sum of squares in tree order (milab200_set_option rmsnorm_fast_reduction; "
                         "not bit-identical to the reference order)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra records (other configs, reference GPU kernels)")
    ap.add_argument("--mode", default="auto", choices=["auto", "chain", "launches"],
                    help="decode (M <= 16): the stack as ONE persistent chained launch (milab200_chain_*) or one launch per Linear")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1 decode: all-reduce fused into the row-parallel GEMV epilogue (NVLink peer memory) or NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
