"""LinearStack — a chain of quantized Linear layers replayed as one CUDA graph.

Mila drives its Linears one `forward` at a time from the transformer block
(Components/Transformers/Gemma/Gemma.Block.ixx:298-349, .../LlaMa/Llama.Block.ixx:883) and lists
CUDA-graph decode as its own next lever (CHANGELOG.md:252-258).  On B200 a decode Linear lasts
1-10 us, below the cost of a host launch, so the B200-idiomatic way to run the path is to capture
the launcher calls once and replay the graph.  This class does only that: it owns no kernels, every
node of the graph is one C-ABI launcher call.

The stack models the Linear part of an MLP block chain (the only non-Linear glue — activation,
norm, residual — is outside this repo's scope), so the data flow is
    h_l --gate_l--> g_l ;  h_l --up_l--> u_l ;  g_l --down_l--> h_{l+1}
i.e. three Linears per layer with the shapes of the model's gate / up / down projections.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib
from .linear import (PerChannelFp8, PerGroupFp4, linear_forward, linear_glu_forward, quantize_fp4_per_group,
                     quantize_fp8_per_channel, rmsnorm_linear_glu_forward)


@dataclass
class QuantWeight:
    weight: torch.Tensor      # uint8 storage
    scales: torch.Tensor      # f32
    N: int
    K: int


def make_quant_weight(N: int, K: int, policy, device, seed: int, row_slice=None, col_slice=None) -> QuantWeight:
    """Random-init (randn / sqrt(K), Linear.ixx:1060) BF16 weights quantized on the device.
    row_slice / col_slice select a tensor-parallel shard of the *unsharded* quantisation
    (FP8 row scales are whole-row absmax), so shards are bit-exact slices."""
    gen = torch.Generator(device=device); gen.manual_seed(seed)
    w = (torch.randn((N, K), device=device, generator=gen, dtype=torch.float32) / K ** 0.5).to(torch.bfloat16)
    if isinstance(policy, PerChannelFp8):
        q, s = quantize_fp8_per_channel(w)
        if row_slice is not None: q, s = q[row_slice].contiguous(), s[row_slice].contiguous()
        if col_slice is not None: q = q[:, col_slice].contiguous()
    else:
        g = policy.kQuantizationGroupSize
        q, s = quantize_fp4_per_group(w, g)
        if row_slice is not None: q, s = q[row_slice].contiguous(), s[row_slice].contiguous()
        if col_slice is not None:
            q = q[:, col_slice.start // 2: col_slice.stop // 2].contiguous()
            s = s[:, col_slice.start // g: col_slice.stop // g].contiguous()
    del w
    return QuantWeight(q, s, q.shape[0], (q.shape[1] * 2) if isinstance(policy, PerGroupFp4) else q.shape[1])


class LinearStack:
    """`layers` MLP-shaped triples (gate, up: hidden->ffn ; down: ffn->hidden), optionally the
    tensor-parallel shard `rank` of `world` (gate/up column-parallel, down row-parallel followed by
    an all-reduce over `group`)."""

    def __init__(self, hidden: int, ffn: int, layers: int, policy, M: int, device="cuda:0",
                 rank: int = 0, world: int = 1, group=None, seed: int = 1234, allreduce: str = "fused",
                 mode: str = "launches", fuse_gate_up: bool = False, glu_kind: int = 2):
        """mode "launches": one C-ABI launcher call per Linear (what Mila's op table does), replayed as a CUDA graph;
        mode "chain": the same Linears as ONE persistent launch (milab200_chain_*, M <= 16).
        fuse_gate_up: gate and up are ONE Linear [2 ffn, hidden] whose epilogue applies the GLU (`glu_kind`, SwiGLU by
        default) — Mila's own MLP dataflow (fc_gate_up + activation, Llama.Block.ixx:883, Gemma.Block.ixx:347): two
        launches per layer instead of three, the same weight bytes.  With per-Linear launches the fused form is the whole
        MLP block as Mila runs it — RMSNorm (ln_2) -> fc_gate_up -> GLU in ONE launch (milab200_rmsnorm_*_gemm_glu), then
        fc_down: the norm also keeps the activations of a deep random-init stack at unit scale (a gated product without it
        collapses to zero or overflows within a few layers)."""
        self.hidden, self.ffn, self.layers, self.policy, self.M = hidden, ffn, layers, policy, M
        self.mode, self.fuse_gate_up, self.glu_kind = mode, fuse_gate_up, glu_kind
        self.norm = bool(fuse_gate_up and mode == "launches")
        # Gemma: (1 + w) with small w (Gemma.Block.ixx:18-20); Llama: w around 1
        self.norm_offset = 1.0 if glu_kind == 1 else 0.0
        self.norm_eps = 1e-6
        if mode not in ("launches", "chain"):
            raise _lib.InvalidArgument(f"LinearStack: unknown mode {mode!r}")
        if mode == "chain" and (M > 16 or (world > 1 and allreduce != "fused")):
            raise _lib.InvalidArgument("LinearStack: chain mode takes M <= 16 and the fused all-reduce")
        self.device = torch.device(device)
        self.rank, self.world, self.group = rank, world, group
        # "fused": decode all-reduce inside the row-parallel GEMV epilogue over NVLink peer memory (csrc/tp.cu);
        # "nccl": ncclAllReduce on the same stream after the row-parallel layer (always used for M > 16)
        self.allreduce = allreduce
        self.tp = None
        if world > 1:
            from .tp import TpGroup
            with torch.cuda.device(torch.device(device)):
                self.tp = TpGroup(group, max_out_features=hidden, device=device)
        assert ffn % world == 0
        shard = ffn // world
        if isinstance(policy, PerGroupFp4):
            assert shard % policy.kQuantizationGroupSize == 0, "row-parallel shard must hold whole groups"
        rs = slice(rank * shard, (rank + 1) * shard) if world > 1 else None
        self.w = []
        with torch.cuda.device(self.device):
            for l in range(layers):
                gate = make_quant_weight(ffn, hidden, policy, self.device, seed + 3 * l, row_slice=rs)
                up = make_quant_weight(ffn, hidden, policy, self.device, seed + 3 * l + 1, row_slice=rs)
                down = make_quant_weight(hidden, ffn, policy, self.device, seed + 3 * l + 2, col_slice=rs)
                if fuse_gate_up:
                    # one Linear: rows [0, shard) gate, [shard, 2 shard) up (the layout fc_gate_up has in Mila)
                    gu = QuantWeight(torch.cat((gate.weight, up.weight), 0), torch.cat((gate.scales, up.scales), 0),
                                     2 * gate.N, gate.K)
                    del gate, up
                    self.w.append((gu, down))
                else:
                    self.w.append((gate, up, down))
            self.h = [torch.zeros((M, hidden), dtype=torch.bfloat16, device=self.device) for _ in range(2)]
            self.g = torch.zeros((M, shard), dtype=torch.bfloat16, device=self.device)
            self.u = torch.zeros((M, shard), dtype=torch.bfloat16, device=self.device)
            self.gu = torch.zeros((M, 2 * shard), dtype=torch.bfloat16, device=self.device) if fuse_gate_up else None
            self.norm_w, self.norm_scratch = [], None
            if self.norm:
                gen = torch.Generator(device=self.device); gen.manual_seed(seed + 7919)
                for l in range(layers):
                    w = 0.1 * torch.randn((hidden,), device=self.device, generator=gen, dtype=torch.float32) + (1.0 - self.norm_offset)
                    self.norm_w.append(w.to(torch.bfloat16))
                self.norm_scratch = torch.zeros((M, hidden), dtype=torch.bfloat16, device=self.device)
        self.chain = None
        if mode == "chain":
            entries, cur = [], 0
            for t in self.w:
                hin, hout = self.h[cur], self.h[cur ^ 1]
                if fuse_gate_up:
                    gu, down = t
                    entries.append({"x": hin, "weight": gu.weight, "scales": gu.scales, "out": self.g, "glu": glu_kind})
                else:
                    gate, up, down = t
                    entries.append({"x": hin, "weight": gate.weight, "scales": gate.scales, "out": self.g})
                    # up reads the same input as gate and nothing gate writes: it may start as soon as gate's own
                    # producer has finished (down, two entries before up, waits for up — so the `u` buffer is never rewritten early)
                    entries.append({"x": hin, "weight": up.weight, "scales": up.scales, "out": self.u,
                                    "depends_on": len(entries) - 2})
                # down reads gate's output only: with separate gate / up entries it need not wait for up (its output buffer
                # `hout` was last read two layers ago; an entry being complete implies every earlier entry is)
                entries.append({"x": self.g, "weight": down.weight, "scales": down.scales, "out": hout, "tp": self.tp,
                                "depends_on": len(entries) - (1 if fuse_gate_up else 2)})
                cur ^= 1
            self._chain_out = self.h[cur]
            self.chain = DecodeChain(entries, policy, M, self.device)
        self.graph: torch.cuda.CUDAGraph | None = None
        self.graph_host: torch.cuda.CUDAGraph | None = None
        self.launches_per_step = 0
        self.x_host = torch.zeros((M, hidden), dtype=torch.bfloat16).pin_memory()
        self.y_host = torch.zeros((M, hidden), dtype=torch.bfloat16).pin_memory()

    # -- bytes / flops bookkeeping (SURVEY.md §8d formulas) ---------------------------------
    def algorithmic_bytes_per_step(self) -> int:
        tot = 0
        for t in self.w:
            for qw in t:
                n_out = qw.N // 2 if (self.fuse_gate_up and qw is t[0]) else qw.N      # the fused GLU writes [M, H]
                tot += qw.weight.numel() + qw.scales.numel() * 4 + 2 * self.M * (qw.K + n_out)
        return tot + sum(2 * w.numel() for w in self.norm_w)

    def weight_bytes(self) -> int:
        return sum(qw.weight.numel() + qw.scales.numel() * 4 for t in self.w for qw in t)

    # -- execution ---------------------------------------------------------------------------
    def _forward_eager(self) -> torch.Tensor:
        if self.chain is not None:
            self.chain.forward()
            return self._chain_out
        cur = 0
        for l, t in enumerate(self.w):
            hin, hout = self.h[cur], self.h[cur ^ 1]
            down = t[-1]
            if self.norm:
                rmsnorm_linear_glu_forward(hin, self.norm_w[l], None, self.norm_eps, self.norm_offset, t[0].weight, t[0].scales,
                                           self.policy, self.glu_kind, None, self.g, self.gu, self.norm_scratch)
            elif self.fuse_gate_up:
                linear_glu_forward(hin, t[0].weight, t[0].scales, self.policy, self.glu_kind, None, self.g, self.gu)
            else:
                gate, up, _ = t
                linear_forward(hin, gate.weight, gate.scales, self.policy, None, self.g)
                linear_forward(hin, up.weight, up.scales, self.policy, None, self.u)
            if self.world > 1:
                self.tp.rowparallel_forward(self.g, down.weight, down.scales, self.policy, None, hout,
                                            force_nccl=(self.allreduce == "nccl"))
            else:
                linear_forward(self.g, down.weight, down.scales, self.policy, None, hout)
            cur ^= 1
        return self.h[cur]

    def capture(self) -> None:
        """Warm up (function attributes, NCCL) on a side stream, then capture one step."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._forward_eager()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            before = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._out = self._forward_eager()
            self.launches_per_step = _lib.launch_count() - before
            # the host-buffer pass as ONE graph launch as well: pinned H2D of the activations, the same launches, D2H of the
            # result (two memcpy nodes around the kernel nodes instead of two cudaMemcpyAsync calls + a graph launch per step)
            self.graph_host = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_host):
                self.h[0].copy_(self.x_host, non_blocking=True)
                out = self._forward_eager()
                self.y_host.copy_(out, non_blocking=True)

    def set_input(self, x: torch.Tensor) -> None:
        self.h[0].copy_(x)

    def step(self) -> torch.Tensor:
        """One pass of the stack with inputs already resident in HBM."""
        if self.graph is None:
            return self._forward_eager()
        self.graph.replay()
        return self._out

    def forward_host(self, x_host: torch.Tensor | None = None) -> torch.Tensor:
        """One pass with HOST buffers: pinned H2D of the step's activations, the stack, D2H of the
        result.  Asynchronous on the current stream; the caller synchronises (as Mila's callers do)."""
        if self.graph_host is not None:
            if x_host is not None and x_host is not self.x_host:
                self.x_host.copy_(x_host)                     # (the graph reads the stack's own pinned staging buffer)
            self.graph_host.replay()
            return self.y_host
        src = self.x_host if x_host is None else x_host
        self.h[0].copy_(src, non_blocking=True)
        out = self.step()
        self.y_host.copy_(out, non_blocking=True)
        return self.y_host


class LayerChain:
    """A chain of arbitrary quantized Linears replayed as one CUDA graph (single GPU): `shapes` lists (K, N) of the
    Linears of ONE transformer layer in call order, repeated `layers` times with distinct weights.  A Linear reads
    the previous Linear's output when the widths match, else a fixed pre-filled buffer of its input width (the
    attention / activation glue between them is outside this repo's scope).  Used to time a whole layer's decode
    Linears back to back, e.g. Gemma 4 12B: qkv 3840->8192, o_proj 4096->3840, fc_gate_up 3840->30720,
    fc_down 15360->3840 (Gemma.Block.ixx:892-918)."""

    def __init__(self, shapes, layers: int, policy, M: int, device="cuda:0", seed: int = 1234):
        self.shapes, self.layers, self.policy, self.M = list(shapes), layers, policy, M
        self.device = torch.device(device)
        self.w = []
        with torch.cuda.device(self.device):
            for l in range(layers):
                self.w.append([make_quant_weight(N, K, policy, self.device, seed + len(self.shapes) * l + i)
                               for i, (K, N) in enumerate(self.shapes)])
            gen = torch.Generator(device=self.device); gen.manual_seed(99)
            self.inputs = {K: torch.randn((M, K), device=self.device, generator=gen).to(torch.bfloat16)
                           for (K, _) in self.shapes}
            self.outputs = [torch.zeros((M, N), dtype=torch.bfloat16, device=self.device) for (_, N) in self.shapes]
        self.graph = None
        self.launches_per_step = 0

    def algorithmic_bytes_per_step(self) -> int:
        return sum(qw.weight.numel() + qw.scales.numel() * 4 + 2 * self.M * (qw.K + qw.N) for layer in self.w for qw in layer)

    def weight_bytes(self) -> int:
        return sum(qw.weight.numel() + qw.scales.numel() * 4 for layer in self.w for qw in layer)

    def _forward_eager(self):
        prev = None
        for layer in self.w:
            for i, qw in enumerate(layer):
                src = prev if (prev is not None and prev.shape[-1] == qw.K) else self.inputs[qw.K]
                prev = linear_forward(src, qw.weight, qw.scales, self.policy, None, self.outputs[i])
        return prev

    def capture(self) -> None:
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._forward_eager()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            before = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._out = self._forward_eager()
            self.launches_per_step = _lib.launch_count() - before

    def step(self):
        self.graph.replay()
        return self._out


class DecodeChain:
    """milab200_chain_*: a list of dependent decode Linears (M <= 16) as ONE persistent launch (csrc/decode_chain.cu).
    `entries`: dicts with x, weight, scales, out and optionally bias, glu (0 / GLU_GEGLU_TANH / GLU_SWIGLU), depends_on
    (default: the previous entry), tp (a TpGroup: row-parallel shard, summed over the ranks in the epilogue).  The
    tensors must stay alive and in place for the life of the chain (their addresses are baked into it)."""

    def __init__(self, entries, policy, M: int, device):
        import ctypes
        from ._lib import ChainLinear
        self.device = torch.device(device)
        self._keep = entries
        arr = (ChainLinear * len(entries))()
        g = 0 if isinstance(policy, PerChannelFp8) else policy.kQuantizationGroupSize
        for i, e in enumerate(entries):
            w = e["weight"]
            K = e["x"].shape[-1]
            arr[i].out_bf16 = e["out"].data_ptr(); arr[i].act_bf16 = e["x"].data_ptr()
            arr[i].weight = w.data_ptr(); arr[i].scales = e["scales"].data_ptr()
            arr[i].bias_bf16 = e["bias"].data_ptr() if e.get("bias") is not None else None
            arr[i].in_features = K; arr[i].out_features = w.shape[0]
            arr[i].group_size = g; arr[i].glu = int(e.get("glu", 0)); arr[i].depends_on = int(e.get("depends_on", i - 1))
            tp = e.get("tp")
            arr[i].tp_ctx = tp._ctx if (tp is not None and tp.world > 1) else None
        ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().milab200_chain_create(arr, len(entries), M, ctypes.byref(ctx)), "chain_create")
        self._ctx = ctx
        self.count = len(entries)

    def forward(self) -> None:
        import ctypes
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(_lib.lib().milab200_chain_forward(self._ctx, st), "chain_forward")

    def describe(self, i: int):
        import ctypes
        r, p, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().milab200_chain_describe(self._ctx, i, ctypes.byref(r), ctypes.byref(p), ctypes.byref(t)), "chain_describe")
        return {"tile_rows": r.value, "ksplits": p.value, "tiles": t.value}

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            _lib.lib().milab200_chain_destroy(self._ctx)
            self._ctx = None
