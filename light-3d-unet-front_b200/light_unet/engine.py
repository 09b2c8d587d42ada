"""Launch plan of the B200-native U-Net: sequences the libl3d kernels for one
forward (and backward) pass of Lightweight3DUNet.

Host-side only: shape bookkeeping, workspace (HBM) layout and kernel ordering.
All arithmetic happens in libl3d.so; there is no PyTorch fallback.

HBM layout (per workspace, i.e. per (N, D, H, W, dtype, training) key):
  * activations are channels-last NDHWC, fp16 by default (fp32 optional);
  * the output of the three encoder blocks that feed a skip connection is
    written straight into the upper channel half of the decoder's concat buffer
    (`cat`), the transposed conv writes the lower half -> torch.cat / F.pad of
    unet3d.py:130-141 never materialise;
  * each block keeps three raw (pre-norm) tensors t1, t2, r plus double
    {sum, sumsq} statistics; norm/LeakyReLU/dropout are applied by the consumer;
  * in training mode the depthwise outputs u1, u2 are also kept for wgrad.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as nv

LEAKY_SLOPE = 0.01   # unet3d.py:52,63
IN_EPS = 1e-5        # nn.InstanceNorm3d default


@dataclass
class BlockSpec:
    name: str          # init_conv, down1, ...
    prefix: str        # state_dict prefix of the residual block
    cin: int
    cout: int
    level: int
    kind1: str         # "dws" | "grouped" | "dense"
    kind2: str


def _conv_kind(dws: bool, grouped: bool, groups: int, cin: int, cout: int, which: int) -> str:
    """Branch selection of ResidualBlock.__init__ (unet3d.py:44-49, :55-60)."""
    if dws:
        return "dws"
    if which == 1:
        ok = grouped and groups > 1 and cin >= groups and cout >= groups
    else:
        ok = grouped and groups > 1 and cout >= groups
    return "grouped" if ok else "dense"


def make_block_specs(in_channels: int, enc: Sequence[int], dws: bool, grouped: bool, groups: int) -> List[BlockSpec]:
    e = list(enc)
    rows = [("init_conv", "init_conv", in_channels, e[0], 0, False),
            ("down1", "down1.res_block", e[0], e[1], 1, grouped),
            ("down2", "down2.res_block", e[1], e[2], 2, grouped),
            ("down3", "down3.res_block", e[2], e[3], 3, grouped),
            ("bottleneck", "bottleneck", e[3], e[3], 3, grouped),
            ("up1", "up1.res_block", e[3], e[2], 2, grouped),
            ("up2", "up2.res_block", e[2], e[1], 1, grouped),
            ("up3", "up3.res_block", e[1], e[0], 0, grouped)]
    return [BlockSpec(n, p, ci, co, lv, _conv_kind(dws, g, groups, ci, co, 1), _conv_kind(dws, g, groups, co, co, 2))
            for n, p, ci, co, lv, g in rows]


class Workspace:
    """All HBM buffers of one forward(/backward) pass for a fixed problem shape."""

    def __init__(self, plan: "UNetPlan", N: int, dims: Tuple[int, int, int], dtype: torch.dtype, device,
                 training: bool):
        self.N, self.dims, self.dtype, self.device, self.training = N, dims, dtype, device, training
        self.cap = N           # samples the buffers were allocated for; an inference workspace is reused for any N <= cap
        e = plan.enc
        lv = [tuple(dims)]
        for _ in range(3):
            lv.append(tuple(d // 2 for d in lv[-1]))
        if min(lv[3]) < 1:
            raise ValueError(f"input spatial size {dims} is too small for three 2x poolings")
        self.level_dims = lv
        z = lambda *s: torch.zeros(*s, dtype=dtype, device=device)
        # L3D_DEBUG_POISON=1 (development aid): every workspace buffer starts as NaN, so a read of a never-written element shows up
        poison = os.environ.get("L3D_DEBUG_POISON", "0") == "1"
        emp = (lambda *s: torch.full(s, float("nan"), dtype=dtype, device=device)) if poison else (lambda *s: torch.empty(*s, dtype=dtype, device=device))
        # concat buffers [up | skip]; zero-initialised once: the centre-pad rim (odd sizes) stays zero
        self.cat = {1: z(N, *lv[1], 2 * e[1]), 2: z(N, *lv[2], 2 * e[2])}
        self.cat0_split = plan.split_cat0(dtype, training)
        if self.cat0_split:
            self.cat0_lo, self.cat0_hi = z(N, *lv[0], e[0]), z(N, *lv[0], e[0])      # [up] and [skip] as two dense tensors
        else:
            self.cat[0] = z(N, *lv[0], 2 * e[0])
        self.pooled = {0: emp(N, *lv[1], e[0]), 1: emp(N, *lv[2], e[1]), 2: emp(N, *lv[3], e[2])}
        self.blocks: Dict[str, dict] = {}
        n_stats = 0
        for b in plan.blocks:
            d = lv[b.level]
            # first block of a 1-channel image at inference: conv1's 16-channel output and the shortcut are rank-1 maps of
            # single-channel tensors and are never stored (UNetPlan.forward); only the depthwise output u (fp32) is
            if b.name == "init_conv" and plan.rank1_first(b, dtype, d, training):
                buf = {"u_r1": torch.empty(N, *d, dtype=torch.float32, device=device), "t2": emp(N, *d, b.cout)}
            else:
                buf = {"t1": emp(N, *d, b.cout), "t2": emp(N, *d, b.cout)}
                if b.cin != b.cout and not plan.rank1_shortcut(b, training):
                    buf["r"] = emp(N, *d, b.cout)
            if training and b.kind1 == "dws":
                buf["u1"] = emp(N, *d, b.cin)
            if training and b.kind2 == "dws":
                buf["u2"] = emp(N, *d, b.cout)
            if b.name in ("down3", "bottleneck", "up1", "up2") or (b.name == "up3" and training):
                buf["out"] = emp(N, *d, b.cout)
            buf["stats_off"] = n_stats
            n_stats += 3 * 2 * N * b.cout
            self.blocks[b.name] = buf
        self.stats = torch.zeros(n_stats, dtype=torch.float64, device=device)
        if training:
            # backward: {sum g, sum g*xhat} per norm (same layout as stats), gradient tensors per level / block
            self.red = torch.zeros(n_stats, dtype=torch.float64, device=device)
            cmax = [2 * e[0], 2 * e[1], 2 * e[2], e[3]]
            # gradient tensors are fp32 whatever the activation storage type: the InstanceNorm backward subtracts
            # the per-(n,c) mean of the incoming gradient, and the Focal Tversky gradient is almost constant over
            # the voxels, so 16-bit rounding of the gradient (relative to its magnitude) swamps the centred signal
            emp = (lambda *s: torch.full(s, float("nan"), dtype=torch.float32, device=device)) if poison else (lambda *s: torch.empty(*s, dtype=torch.float32, device=device))
            self.g_cat = {k: emp(N, *lv[k], 2 * e[k]) for k in range(3)}      # grad of [up | skip]
            self.g_pooled = {k: emp(N, *lv[k + 1], e[k]) for k in range(3)}   # grad of the pooled block inputs
            self.g_out = {b.name: emp(N, *lv[b.level], b.cout) for b in plan.blocks
                          if b.name in ("down3", "bottleneck", "up1", "up2")}
            self.gz = {k: emp(N, *lv[k], e[k]) for k in range(4)}             # grad of the pre-activation merge sum
            self.gy = {k: emp(N, *lv[k], e[k]) for k in range(4)}             # grad of norm1's output
            self.gu = {k: emp(N, *lv[k], cmax[k]) for k in range(4)}          # grad of a depthwise output
            self._conv3_work = None
        self.generation = 0
        self.prob_out = None
        self.logits = None
        self.nbytes = sum(t.numel() * t.element_size() for t in self._tensors())

    def _tensors(self):
        for v in vars(self).values():
            if isinstance(v, torch.Tensor):
                yield v
            elif isinstance(v, dict):
                for w in v.values():
                    if isinstance(w, torch.Tensor):
                        yield w
                    elif isinstance(w, dict):
                        yield from (x for x in w.values() if isinstance(x, torch.Tensor))

    def stats_of(self, name: str, which: int, cout: int) -> torch.Tensor:
        off = self.blocks[name]["stats_off"] + which * 2 * self.N * cout
        return self.stats[off: off + 2 * self.N * cout]

    def red_of(self, name: str, which: int, cout: int) -> torch.Tensor:
        off = self.blocks[name]["stats_off"] + which * 2 * self.N * cout
        return self.red[off: off + 2 * self.N * cout]

    def conv3_work(self, plan: "UNetPlan") -> torch.Tensor:
        """Scratch of the dense / grouped 3x3x3 backward: two activation-sized tensors at the widest level plus
        the repacked weights (see l3d_conv3_bwd_workspace_bytes)."""
        if self._conv3_work is None:
            need = 0
            for b in plan.blocks:
                d = self.level_dims[b.level]
                need = max(need, int(nv.lib().l3d_conv3_bwd_workspace_bytes(self.N, d[0], d[1], d[2], max(b.cin, b.cout), b.cout,
                                                                            4)))
            self._conv3_work = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._conv3_work


class UNetPlan:
    def __init__(self, in_channels: int, out_channels: int, enc: Sequence[int], dws: bool, grouped: bool, groups: int):
        if len(enc) != 4:
            raise ValueError("encoder_channels must have 4 entries")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.enc = list(enc)
        self.dws, self.grouped, self.groups = dws, grouped, groups
        self.blocks = make_block_specs(in_channels, enc, dws, grouped, groups)
        self._ws: Dict[tuple, Workspace] = {}

    # first block of a 1-channel image at inference (b.cin == 1, depthwise-separable): the shortcut r = sc (x) x and
    # conv1's output t1 = pw (x) u are rank-1 maps of single-channel tensors, evaluated on the fly by their consumers
    def rank1_shortcut(self, b: BlockSpec, training: bool) -> bool:
        return (not training) and b.name == "init_conv" and b.cin == 1 and b.cin != b.cout and b.kind1 == "dws" and b.cout in (16, 32)

    def split_cat0(self, dtype, training: bool) -> bool:
        """Inference, fp16 storage, 16 + 16 channels at the top level, depthwise-separable up3.conv1: the decoder block reads
        [ConvTranspose output | skip] as two dense 16-channel tensors (l3d_dwpw_fwd2) instead of one interleaved buffer."""
        up3 = self.blocks[-1]
        return ((not training) and dtype == torch.float16 and up3.kind1 == "dws" and up3.cin == 32 and up3.cout == 16 and self.enc[0] == 16
                and os.environ.get("L3D_SPLIT_CAT", "1") != "0" and os.environ.get("L3D_NO_IGEMM", "0") != "1")   # needs the implicit-GEMM kernel

    def rank1_first(self, b: BlockSpec, dtype, dims, training: bool) -> bool:
        return (self.rank1_shortcut(b, training) and b.kind2 == "dws" and b.cout == 16 and dtype == torch.float16
                and dims[2] % 4 == 0 and os.environ.get("L3D_NO_RANK1_FIRST", "0") != "1")

    # ------------------------------------------------------------------ workspace
    # Cache policy: ONE inference workspace per (dims, dtype, device), allocated for the largest batch seen and reused for
    # every smaller one (a sliding-window tail batch, a volume with fewer windows); training workspaces are exact-N (the
    # backward pass reads them) and at most MAX_TRAIN_WS are kept.  A new allocation first drops whatever would push the
    # cached total past WS_BUDGET_FRAC of the device memory (a 325-window fp16 workspace is ~20 GB).
    MAX_TRAIN_WS = 2
    MAX_INFER_WS = 2
    WS_BUDGET_FRAC = 0.3

    def inference_capacity(self, dims, dtype, device) -> int:
        ws = self._ws.get((tuple(dims), dtype, str(device), False))
        return ws.cap if ws is not None else 0

    def cached_bytes(self) -> int:
        return sum(w.nbytes for w in self._ws.values())

    def drop_workspaces(self, training=None):
        for k in [k for k in self._ws if training is None or k[-1] == training]:
            del self._ws[k]

    def workspace(self, N, dims, dtype, device, training) -> Workspace:
        dims = tuple(dims)
        key = (N, dims, dtype, str(device), True) if training else (dims, dtype, str(device), False)
        ws = self._ws.pop(key, None)
        if ws is not None and ws.cap < N:
            ws = None                               # grow: the old buffers are released before the new ones are allocated
        if ws is None:
            same = [k for k in self._ws if k[-1] == training]
            limit = (self.MAX_TRAIN_WS if training else self.MAX_INFER_WS) - 1
            for k in same[:max(0, len(same) - limit)]:         # dicts keep insertion order: oldest first
                del self._ws[k]
            if self._ws and torch.device(device).type == "cuda":
                total = torch.cuda.get_device_properties(device).total_memory
                if self.cached_bytes() > self.WS_BUDGET_FRAC * total:
                    self._ws.clear()
            ws = Workspace(self, N, dims, dtype, device, training)
        ws.N = N
        self._ws[key] = ws                          # most recently used last
        return ws

    # -------------------------------------------------------------------- forward
    def _conv(self, P, b: BlockSpec, which: int, x_act, xn, N, dims, t, t_stats, sc_w, r, r_stats, u, st):
        """conv1 / conv2 of a residual block (+ fused shortcut where the kernel supports it)."""
        kind = b.kind1 if which == 1 else b.kind2
        pre = f"{b.prefix}.conv{which}"
        nv.TIMER.tag = f"{b.name}.c{which}"
        D, H, W = dims
        es, nvx = t.element_size(), N * D * H * W
        cin, cout = x_act.C, t.shape[-1]
        if kind == "dws":
            nv.call("l3d_dwpw_fwd", x_act, xn, N, D, H, W, nv.ptr(P[f"{pre}.depthwise.weight"]),
                    nv.ptr(P[f"{pre}.pointwise.weight"]), nv.ptr(sc_w), nv.act(t), nv.ptr(t_stats),
                    nv.act(r), nv.ptr(r_stats), nv.act(u), st,
                    algo_bytes=es * nvx * (cin + cout * (2 if (sc_w is not None and r is not None) else 1) + (cin if u is not None else 0)))
        else:
            w = P[f"{pre}.conv.weight"] if kind == "grouped" else P[f"{pre}.weight"]
            g = self.groups if kind == "grouped" else 1
            nv.call("l3d_conv3_fwd", x_act, xn, N, D, H, W, nv.ptr(w), g, nv.act(t), nv.ptr(t_stats),
                    nv.ptr(sc_w), nv.act(r), nv.ptr(r_stats), st,
                    algo_bytes=es * nvx * (cin + cout * (2 if sc_w is not None else 1)))

    def forward(self, P: Dict[str, torch.Tensor], x_cl: torch.Tensor, training: bool,
                masks: Optional[List[Optional[torch.Tensor]]] = None,
                prob_out: Optional[torch.Tensor] = None) -> Workspace:
        """x_cl: [N, D, H, W, Cin] channels-last activation-dtype tensor.  Returns the workspace holding
        `prob` ([N, OC, D, H, W] fp32) and, in training mode, everything the backward pass needs."""
        nv.require_cuda(x_cl, "UNetPlan.forward")
        e = self.enc
        for k in range(3):
            # UpBlock: cat([ConvTranspose3d(C, C/2)(x), skip]) feeds ResidualBlock(C, ...) (unet3d.py:119-141): the reference
            # fails inside that conv when the widths do not add up; say so up front
            if e[k + 1] // 2 + e[k] != e[k + 1]:
                raise RuntimeError(f"encoder_channels {e}: up-block {3 - k} concatenates {e[k + 1] // 2} + {e[k]} channels "
                                   f"but its residual block expects {e[k + 1]} (each width must be twice the previous one)")
        N, D, H, W, _ = x_cl.shape
        ws = self.workspace(N, (D, H, W), x_cl.dtype, x_cl.device, training)
        ws.x = x_cl
        ws.masks = masks
        ws.generation += 1
        # fresh output tensors every call: the caller owns them (workspace buffers are reused by the next forward)
        if prob_out is None:
            prob_out = torch.empty(N, self.out_channels, D, H, W, dtype=torch.float32, device=x_cl.device)
        assert prob_out.is_contiguous() and prob_out.dtype == torch.float32 and prob_out.numel() == N * self.out_channels * D * H * W
        ws.prob_out = prob_out
        ws.logits = torch.empty_like(ws.prob_out) if training else None
        ws.stats.zero_()
        st = nv.stream_ptr(x_cl.device)
        ident = nv.norm()
        cur = x_cl                       # materialised activation feeding the next block
        cur_off, cur_C = 0, x_cl.shape[-1]
        for i, b in enumerate(self.blocks):
            nv.TIMER.tag = b.name
            dims = ws.level_dims[b.level]
            vox = dims[0] * dims[1] * dims[2]
            es = x_cl.element_size()
            mbytes = es * N * vox * b.cout            # one tensor of this block's output shape
            buf = ws.blocks[b.name]
            mask = masks[i] if (masks is not None and masks[i] is not None) else None
            if b.name.startswith("up"):
                # transposed conv into the lower half of the concat buffer, then the block reads the whole buffer
                split = b.level == 0 and ws.cat0_split
                cat = None if split else ws.cat[b.level]
                lo = ws.level_dims[b.level + 1]
                off = [(dims[k] - 2 * lo[k]) // 2 for k in range(3)]
                nv.call("l3d_convt_fwd", nv.act(cur, cur_off, cur_C), N, lo[0], lo[1], lo[2],
                        nv.ptr(P[f"{b.name}.up.weight"]), nv.ptr(P[f"{b.name}.up.bias"]),
                        nv.act(ws.cat0_lo) if split else nv.act(cat, 0, b.cin // 2), dims[0], dims[1], dims[2], off[0], off[1], off[2], st,
                        algo_bytes=es * N * (lo[0] * lo[1] * lo[2] * cur_C + vox * (b.cin // 2)))
                cur, cur_off, cur_C = (ws.cat0_lo, 0, b.cin // 2) if split else (cat, 0, b.cin)
            split_in = b.name == "up3" and ws.cat0_split
            x_act = nv.act(cur, cur_off, cur_C)
            has_sc = b.cin != b.cout
            s1, s2, sr = (ws.stats_of(b.name, k, b.cout) for k in range(3))
            sc_w = P[f"{b.prefix}.shortcut.0.weight"] if has_sc else None
            # inference, first block of a 1-channel image: its 1x1x1 shortcut r[v][c] = w[c] * x[v] is never written -- the
            # conv kernel still produces r's statistics and the merge evaluates r on the fly (64 B / voxel less traffic)
            rank1 = self.rank1_shortcut(b, training)
            n1 = nv.norm(s1, P[f"{b.prefix}.norm1.weight"], P[f"{b.prefix}.norm1.bias"], mask, IN_EPS, LEAKY_SLOPE, vox)
            if "u_r1" in buf:
                # ... and conv1's output t1[v][c] = pw[c] * u[v] is not written either: one fp32 channel u plus analytic
                # statistics, then conv2 evaluates lrelu(IN1(t1)) from u on the fly
                pw1 = P[f"{b.prefix}.conv1.pointwise.weight"]
                nv.TIMER.tag = f"{b.name}.c1"
                nv.call("l3d_dw_c1_fwd", x_act, ident, N, *dims, nv.ptr(P[f"{b.prefix}.conv1.depthwise.weight"]), nv.ptr(pw1),
                        nv.ptr(sc_w), b.cout, nv.ptr(buf["u_r1"]), nv.ptr(s1), nv.ptr(sr), st,
                        algo_bytes=N * vox * (es + 4))
                nv.TIMER.tag = f"{b.name}.c2"
                nv.call("l3d_dwpw_fwd_rank1", nv.ptr(buf["u_r1"]), nv.ptr(pw1), b.cout, n1, N, *dims,
                        nv.ptr(P[f"{b.prefix}.conv2.depthwise.weight"]), nv.ptr(P[f"{b.prefix}.conv2.pointwise.weight"]),
                        nv.act(buf["t2"]), nv.ptr(s2), st, algo_bytes=N * vox * (4 + es * b.cout))
            else:
                # conv1 (+ shortcut conv) on the block input
                if split_in:
                    pre = f"{b.prefix}.conv1"
                    nv.TIMER.tag = f"{b.name}.c1"
                    nv.call("l3d_dwpw_fwd2", x_act, nv.act(ws.cat0_hi), ident, N, *dims, nv.ptr(P[f"{pre}.depthwise.weight"]),
                            nv.ptr(P[f"{pre}.pointwise.weight"]), nv.ptr(sc_w), nv.act(buf["t1"]), nv.ptr(s1), nv.act(buf["r"]), nv.ptr(sr), st,
                            algo_bytes=es * N * vox * (b.cin + 2 * b.cout))
                else:
                    self._conv(P, b, 1, x_act, ident, N, dims, buf["t1"], s1, sc_w, None if rank1 else buf.get("r"), sr if has_sc else None,
                               buf.get("u1"), st)
                # conv2 on lrelu(IN1(t1)) * dropout-mask, applied on load
                self._conv(P, b, 2, nv.act(buf["t1"]), n1, N, dims, buf["t2"], s2, None, None, None, buf.get("u2"), st)
            # residual merge (+ pool / head)
            nv.TIMER.tag = b.name
            n2 = nv.norm(s2, P[f"{b.prefix}.norm2.weight"], P[f"{b.prefix}.norm2.bias"], None, IN_EPS, 1.0, vox)
            if has_sc:
                r_act = None if rank1 else nv.act(buf["r"])
                nr = nv.norm(sr, P[f"{b.prefix}.shortcut.1.weight"], P[f"{b.prefix}.shortcut.1.bias"], None, IN_EPS, 1.0, vox)
            else:
                r_act, nr = x_act, ident
            if b.name in ("init_conv", "down1", "down2"):
                # the block output is the skip: upper half of the concat buffer (or the dense skip tensor of the split top level)
                skip_act = nv.act(ws.cat0_hi) if (b.level == 0 and ws.cat0_split) else nv.act(ws.cat[b.level], b.cout, b.cout)
                if rank1:
                    nv.call("l3d_merge_fwd_rank1", nv.act(buf["t2"]), n2, x_act, nv.ptr(sc_w), nr, N, *dims, LEAKY_SLOPE,
                            skip_act, nv.act(ws.pooled[b.level]), st,
                            algo_bytes=2 * mbytes + mbytes // 8 + mbytes // b.cout)
                else:
                    nv.call("l3d_merge_fwd", nv.act(buf["t2"]), n2, r_act, nr, N, *dims, LEAKY_SLOPE,
                            skip_act, nv.act(ws.pooled[b.level]), None, None, 0, None, None, st,
                            algo_bytes=3 * mbytes + mbytes // 8)
                cur, cur_off, cur_C = ws.pooled[b.level], 0, b.cout
            elif b.name == "up3":
                out = buf.get("out")
                nv.call("l3d_merge_fwd", nv.act(buf["t2"]), n2, r_act, nr, N, *dims, LEAKY_SLOPE,
                        nv.act(out), nv.act(None), nv.ptr(P["out_conv.weight"]), nv.ptr(P["out_conv.bias"]),
                        self.out_channels, nv.ptr(ws.prob_out), nv.ptr(ws.logits), st,
                        algo_bytes=(3 if out is not None else 2) * mbytes + 4 * N * vox * self.out_channels * (2 if training else 1))
            else:
                nv.call("l3d_merge_fwd", nv.act(buf["t2"]), n2, r_act, nr, N, *dims, LEAKY_SLOPE,
                        nv.act(buf["out"]), nv.act(None), None, None, 0, None, None, st, algo_bytes=3 * mbytes)
                cur, cur_off, cur_C = buf["out"], 0, b.cout
        return ws

    # ------------------------------------------------------------------- backward
    def _conv_bwd(self, P, G, b: BlockSpec, which: int, ws: Workspace, g, t, nt, red, x_act, xn, u, gy, acc_gy, redx,
                  N, dims, st):
        """Backward of conv1 / conv2: g is the gradient w.r.t. the normalised output of raw tensor t."""
        kind = b.kind1 if which == 1 else b.kind2
        pre = f"{b.prefix}.conv{which}"
        D, H, W = dims
        if kind == "dws":
            cin = x_act.C
            gu = nv.act(ws.gu[b.level], 0, cin)
            nv.call("l3d_pw_bwd", nv.act(g), nv.act(t), nt, nv.ptr(red), nv.act(u), nv.norm(), N, D, H, W,
                    nv.ptr(P[f"{pre}.pointwise.weight"]), nv.ptr(G[f"{pre}.pointwise.weight"]), gu, 0, st)
            nv.call("l3d_dw_bwd", gu, x_act, xn, N, D, H, W, nv.ptr(P[f"{pre}.depthwise.weight"]),
                    nv.ptr(G[f"{pre}.depthwise.weight"]), gy, acc_gy, nv.ptr(redx), st)
        else:
            key = f"{pre}.conv.weight" if kind == "grouped" else f"{pre}.weight"
            g_ = self.groups if kind == "grouped" else 1
            work = ws.conv3_work(self)
            nv.call("l3d_conv3_bwd", nv.act(g), nv.act(t), nt, nv.ptr(red), x_act, xn, N, D, H, W, nv.ptr(P[key]), g_,
                    nv.ptr(G[key]), gy, acc_gy, nv.ptr(redx), nv.ptr(work), work.numel(), st)

    def backward(self, P: Dict[str, torch.Tensor], ws: Workspace, g_prob: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Kernel sequence of the backward pass for the forward recorded in `ws`.  Returns fp32 gradients in the
        parameters' own (PyTorch) layouts, keyed by state_dict name."""
        nv.require_cuda(g_prob, "UNetPlan.backward")
        N = ws.N
        dev = ws.device
        st = nv.stream_ptr(dev)
        # one flat fp32 buffer (a single fill launch) viewed per parameter; autograd takes the views as .grad
        flat = torch.zeros(sum(v.numel() for v in P.values()), dtype=torch.float32, device=dev)
        G, off = {}, 0
        for k, v in P.items():
            G[k] = flat[off:off + v.numel()].view(v.shape)
            off += v.numel()
        ws.red.zero_()
        g_prob = g_prob.to(torch.float32).contiguous()
        ident = nv.norm()
        masks = ws.masks
        nblk = len(self.blocks)
        norm_jobs = []
        for i in range(nblk - 1, -1, -1):
            b = self.blocks[i]
            nv.TIMER.tag = b.name
            dims = ws.level_dims[b.level]
            vox = dims[0] * dims[1] * dims[2]
            buf = ws.blocks[b.name]
            has_sc = b.cin != b.cout
            mask = masks[i] if (masks is not None and masks[i] is not None) else None
            s1, s2, sr = (ws.stats_of(b.name, k, b.cout) for k in range(3))
            r1, r2, rr = (ws.red_of(b.name, k, b.cout) for k in range(3))
            n1 = nv.norm(s1, P[f"{b.prefix}.norm1.weight"], P[f"{b.prefix}.norm1.bias"], mask, IN_EPS, LEAKY_SLOPE, vox)
            n2 = nv.norm(s2, P[f"{b.prefix}.norm2.weight"], P[f"{b.prefix}.norm2.bias"], None, IN_EPS, 1.0, vox)
            # ---- where the block input lives, and where its gradient goes
            if b.name == "init_conv":
                x_act, g_in_t = nv.act(ws.x), None
            elif b.name.startswith("down"):
                x_act, g_in_t = nv.act(ws.pooled[b.level - 1]), ws.g_pooled[b.level - 1]
            elif b.name == "bottleneck":
                x_act, g_in_t = nv.act(ws.blocks["down3"]["out"]), ws.g_out["down3"]
            else:
                x_act, g_in_t = nv.act(ws.cat[b.level]), ws.g_cat[b.level]
            g_in = nv.act(g_in_t)
            if has_sc:
                r_act = nv.act(buf["r"])
                nr = nv.norm(sr, P[f"{b.prefix}.shortcut.1.weight"], P[f"{b.prefix}.shortcut.1.bias"], None, IN_EPS, 1.0, vox)
            else:
                r_act, nr = x_act, ident
            gz = ws.gz[b.level]
            # ---- 1. residual merge (+ pool / head) backward
            if b.name in ("init_conv", "down1", "down2"):
                cat = ws.cat[b.level]
                nv.call("l3d_merge_bwd", nv.act(ws.g_cat[b.level], b.cout, b.cout), nv.act(ws.g_pooled[b.level]),
                        nv.act(cat, b.cout, b.cout), nv.act(ws.pooled[b.level]), nv.act(buf["t2"]), n2, r_act, nr,
                        N, *dims, LEAKY_SLOPE, None, 0, None, None, None, None, nv.act(gz), nv.ptr(r2),
                        nv.ptr(rr) if has_sc else None, st)
            elif b.name == "up3":
                nv.call("l3d_merge_bwd", nv.act(None), nv.act(None), nv.act(buf["out"]), nv.act(None), nv.act(buf["t2"]), n2,
                        r_act, nr, N, *dims, LEAKY_SLOPE, nv.ptr(P["out_conv.weight"]), self.out_channels, nv.ptr(g_prob),
                        nv.ptr(ws.prob_out), nv.ptr(G["out_conv.weight"]), nv.ptr(G["out_conv.bias"]), nv.act(gz),
                        nv.ptr(r2), nv.ptr(rr) if has_sc else None, st)
            else:
                nv.call("l3d_merge_bwd", nv.act(ws.g_out[b.name]), nv.act(None), nv.act(buf["out"]), nv.act(None),
                        nv.act(buf["t2"]), n2, r_act, nr, N, *dims, LEAKY_SLOPE, None, 0, None, None, None, None,
                        nv.act(gz), nv.ptr(r2), nv.ptr(rr) if has_sc else None, st)
            # ---- 2. conv2 backward: gz -> gy (gradient of norm1's output, through dropout / LeakyReLU)
            gy = ws.gy[b.level]
            self._conv_bwd(P, G, b, 2, ws, gz, buf["t2"], n2, r2, nv.act(buf["t1"]), n1, buf.get("u2"), nv.act(gy), 0, r1,
                           N, dims, st)
            # ---- 3. shortcut backward writes g_in, conv1 backward accumulates into it
            need_gin = bool(g_in.ptr)
            if has_sc:
                nv.call("l3d_pw_bwd", nv.act(gz), r_act, nr, nv.ptr(rr), x_act, ident, N, *dims,
                        nv.ptr(P[f"{b.prefix}.shortcut.0.weight"]), nv.ptr(G[f"{b.prefix}.shortcut.0.weight"]),
                        g_in, 0, st)
            elif need_gin:
                # identity shortcut (cin == cout: the bottleneck, or a down block of an encoder with equal adjacent widths):
                # d(out)/d(x) passes gz through into the block's OWN input gradient, which conv1's backward then adds to
                g_in_t[:N].copy_(gz[:N])
            n1b = nv.norm(s1, P[f"{b.prefix}.norm1.weight"], P[f"{b.prefix}.norm1.bias"], None, IN_EPS, 1.0, vox)
            self._conv_bwd(P, G, b, 1, ws, gy, buf["t1"], n1b, r1, x_act, ident, buf.get("u1"), g_in, 1, None,
                           N, dims, st)
            # ---- 4. InstanceNorm affine gradients: collected, one launch after the last block
            norm_jobs.append((r1, b.cout, G[f"{b.prefix}.norm1.weight"], G[f"{b.prefix}.norm1.bias"]))
            norm_jobs.append((r2, b.cout, G[f"{b.prefix}.norm2.weight"], G[f"{b.prefix}.norm2.bias"]))
            if has_sc:
                norm_jobs.append((rr, b.cout, G[f"{b.prefix}.shortcut.1.weight"], G[f"{b.prefix}.shortcut.1.bias"]))
            # ---- 5. transposed conv backward: lower half of g_cat -> gradient of the previous block's output
            if b.name.startswith("up"):
                prev = self.blocks[i - 1]
                lo = ws.level_dims[b.level + 1]
                off = [(dims[k] - 2 * lo[k]) // 2 for k in range(3)]
                nv.call("l3d_convt_bwd", nv.act(ws.g_cat[b.level], 0, b.cin // 2), dims[0], dims[1], dims[2],
                        off[0], off[1], off[2], nv.act(ws.blocks[prev.name]["out"]), N, lo[0], lo[1], lo[2],
                        nv.ptr(P[f"{b.name}.up.weight"]), nv.ptr(G[f"{b.name}.up.weight"]), nv.ptr(G[f"{b.name}.up.bias"]),
                        nv.act(ws.g_out[prev.name]), 0, st)
        nv.TIMER.tag = "norms"
        cnt = len(norm_jobs)
        ptrs = lambda k: (ctypes.c_void_p * cnt)(*[j[k].data_ptr() for j in norm_jobs])
        nv.call("l3d_norm_param_grad_batch", cnt, ptrs(0), (ctypes.c_int * cnt)(*[j[1] for j in norm_jobs]), ptrs(2), ptrs(3), N, st)
        return G
