"""Host-side runtime of the native forward: owns the packed weights, the native plan, the workspace
and (optionally) CUDA graphs.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native
from .pack import pack_state_dict

_TORCH2LMV = {torch.bfloat16: _native.DTYPE_BF16, torch.float32: _native.DTYPE_F32}


def require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            "lemevit_b200 runs on CUDA (sm_100a) only: the forward pass is a set of hand-written Blackwell "
            "kernels and there is deliberately no CPU fallback. Move the model and input to a B200.")


class Engine:
    """One engine per (module, device).  Not thread-safe; one engine per thread/replica."""

    def __init__(self, sd: Dict[str, torch.Tensor], *, depth, embed_dim, mlp_ratios, attn_type, head_dim,
                 queries_len, num_classes, in_chans, backbone: bool, device: torch.device, chunk: int = 0):
        self.lib = _native.load()
        self.device = torch.device(device)
        self.backbone = bool(backbone)
        self.embed_dim = list(embed_dim)
        self.num_classes = int(num_classes)
        self.in_chans = int(in_chans)
        mlp_hidden = [int(r * d) for r, d in zip(mlp_ratios, embed_dim)]
        self.cfg = _native.make_config(depth, embed_dim, mlp_hidden, attn_type, head_dim, queries_len,
                                       num_classes, in_chans, backbone)
        with torch.cuda.device(self.device):
            self.packed: List[torch.Tensor] = pack_state_dict(
                sd, depth=depth, embed_dim=embed_dim, attn_type=attn_type, in_chans=in_chans,
                num_classes=num_classes, backbone=backbone, device=self.device)
        n = len(self.packed)
        arr = (_native.Tensor * n)()
        for i, t in enumerate(self.packed):
            arr[i].data = t.data_ptr()
            arr[i].numel = t.numel()
            arr[i].dtype = _TORCH2LMV[t.dtype]
        self._plan = C.c_void_p()
        _native.check(self.lib.lmv_plan_create(C.byref(self.cfg), arr, n, C.byref(self._plan)))
        self._ws: Optional[torch.Tensor] = None
        # captured graphs: key -> (static_x, static_out, graph, workspaces, packed).  A graph bakes raw device pointers of its
        # workspace and of the packed weights into its kernel nodes, so every entry OWNS its workspace tensors and holds a
        # reference to the packed list: neither can be freed or reallocated while the graph (or a replay callable handed to
        # a caller) is alive.  close() bumps the generation, after which stale replay callables raise instead of running.
        self._graphs: "OrderedDict[Tuple, Tuple]" = OrderedDict()
        self.max_graphs = int(os.environ.get("LEMEVIT_B200_MAX_GRAPHS", "8"))
        self._generation = 0
        self._closed = False
        # concurrent sub-batches per forward: two streams measured +4 % on Base b256 (four: -0.5 %); small batches stay on one
        self.lanes = int(os.environ.get("LEMEVIT_B200_LANES", "2"))
        self.lane_min_batch = 128
        self.chunk = 0
        if chunk:
            self.set_chunk(chunk)

    # -- lifecycle ---------------------------------------------------------------------------------
    def close(self):
        """Destroy the native plan.  Captured graphs die with it: replay callables obtained from graphed() raise afterwards."""
        self._closed = True
        self._invalidate_graphs()
        if getattr(self, "_plan", None) is not None and self._plan.value:
            self.lib.lmv_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def _invalidate_graphs(self):
        self._generation = getattr(self, "_generation", 0) + 1
        if getattr(self, "_graphs", None):
            # a graph that is still replaying must finish before its workspace can go back to the allocator
            torch.cuda.synchronize(self.device)
            self._graphs.clear()

    def _check_open(self):
        if self._closed:
            raise RuntimeError("lemevit_b200: this engine was closed (the module's weights changed or it was re-created); "
                               "call model.native_engine(device) again")

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_chunk(self, images_per_chunk: int):
        _native.check(self.lib.lmv_plan_set_chunk(self._plan, int(images_per_chunk)))
        self.chunk = int(images_per_chunk)
        self._invalidate_graphs()

    def set_debug_simt(self, enable: bool):
        _native.check(self.lib.lmv_plan_set_debug_simt(self._plan, int(bool(enable))))
        self._invalidate_graphs()

    def set_tap(self, stage: int, block: int, x_tokens: Optional[torch.Tensor] = None, c: Optional[torch.Tensor] = None):
        """Test hook: copy (x tokens [B, N, C], c [B, M, C]) after block (stage, block) into the given bf16 buffers on every
        following forward; stage < 0 turns it off.  The caller keeps the buffers alive."""
        _native.check(self.lib.lmv_plan_set_tap(self._plan, int(stage), int(block), x_tokens.data_ptr() if x_tokens is not None else None,
                                                c.data_ptr() if c is not None else None))
        self._invalidate_graphs()

    def set_option(self, name: str, value: int):
        """Schedule A/B switches of the native plan (lmv_plan_set_option), e.g. ``fused_mlp``."""
        _native.check(self.lib.lmv_plan_set_option(self._plan, name.encode(), int(value)))
        self._invalidate_graphs()

    def set_profile(self, enable: bool):
        """Per-kernel-class CUDA-event timing of every following forward (see lmv_plan_set_profile)."""
        _native.check(self.lib.lmv_plan_set_profile(self._plan, int(bool(enable))))

    def get_profile(self):
        arr = (_native.ProfileEntry * 16)()
        n = int(self.lib.lmv_plan_get_profile(self._plan, arr, 16))
        if n < 0:
            _native.check(n)
        return [dict(name=arr[i].name.decode(), launches=arr[i].launches, device_ms=arr[i].device_ms, flops=arr[i].flops,
                     bytes=arr[i].bytes) for i in range(n)]

    def profile_report(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        n = int(self.lib.lmv_plan_profile_report(self._plan, buf, len(buf)))
        if n < 0:
            _native.check(n)
        return buf.value.decode()

    # -- helpers -----------------------------------------------------------------------------------
    def _workspace(self, B: int, H: int, W: int) -> torch.Tensor:
        need = int(self.lib.lmv_workspace_bytes(self._plan, B, H, W))
        if need == 0:
            raise RuntimeError(f"lemevit_b200: unsupported input shape {B}x{H}x{W}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)  # torch OOM -> RuntimeError "out of memory"
        return self._ws

    def _prep_input(self, x: torch.Tensor) -> torch.Tensor:
        """Float input: contiguous NCHW (bf16 / fp32; anything else is cast to bf16).  uint8 input (raw pixels, normalised inside
        the stem kernel, see set_input_norm): [B, 3, H, W] planar (timm fast_collate), the same in channels_last memory, or
        [B, H, W, 3] interleaved — the latter two are handed to the kernel as they lie in memory, seen through an NCHW view."""
        require_cuda(x)
        if x.dtype == torch.uint8:
            if x.dim() != 4 or self.in_chans != 3 or 3 not in (x.shape[1], x.shape[3]):
                raise RuntimeError(f"expected 8-bit input [B, 3, H, W] or [B, H, W, 3] (in_chans = {self.in_chans}), got {tuple(x.shape)}")
            if x.shape[1] != 3:                                   # [B, H, W, 3]
                return x.contiguous().permute(0, 3, 1, 2)
            if x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last):
                return x
            return x.contiguous()
        if x.dim() != 4 or x.shape[1] != self.in_chans:
            raise RuntimeError(f"expected input [B, {self.in_chans}, H, W], got {tuple(x.shape)}")
        if x.dtype not in _TORCH2LMV:
            x = x.to(torch.bfloat16)
        return x.contiguous()      # NCHW; channels_last inputs are re-laid out here (benchmark.py --channels-last)

    @staticmethod
    def _x_code(x: torch.Tensor) -> int:
        """LMV_DTYPE_* of a tensor that went through _prep_input."""
        if x.dtype == torch.uint8:
            return _native.DTYPE_U8 if x.is_contiguous() else _native.DTYPE_U8_NHWC
        return _TORCH2LMV[x.dtype]

    def set_input_norm(self, mean, std):
        """Per-channel mean / std (pixel units, 0..255) of the 8-bit input path: the stem computes
        bf16((float(u8) - mean[c]) / std[c]) per tap, the bits of ``x.float().sub_(mean).div_(std).to(torch.bfloat16)``."""
        m = (C.c_float * 3)(*[float(v) for v in mean])
        s = (C.c_float * 3)(*[float(v) for v in std])
        _native.check(self.lib.lmv_plan_set_input_norm(self._plan, m, s))
        self._invalidate_graphs()

    def launch_count(self, B: int, H: int, W: int) -> int:
        if not self.backbone and self.lanes > 1 and B >= self.lane_min_batch and B % self.lanes == 0:
            return self.lanes * self.launch_count_single(B // self.lanes, H, W)
        return self.launch_count_single(B, H, W)

    def launch_count_single(self, B: int, H: int, W: int) -> int:
        n = int(self.lib.lmv_launch_count(self._plan, B, H, W))
        if n < 0:
            _native.check(n)
        return n

    def out_shapes(self, B: int, H: int, W: int) -> List[Tuple[int, int, int, int]]:
        h, w = (H + 1) // 2, (W + 1) // 2
        h, w = (h + 1) // 2, (w + 1) // 2
        shapes = []
        at = self.cfg.attn_type
        for i in range(self.cfg.num_stages):
            if i > 0 and at[i - 1:i] != b"C":
                h, w = (h + 1) // 2, (w + 1) // 2
            if i >= 1:
                shapes.append((B, self.embed_dim[i], h, w))
        return shapes

    # -- forward -----------------------------------------------------------------------------------
    def _use_lanes(self, B: int) -> int:
        return self.lanes if (not self.backbone and self.lanes > 1 and B >= self.lane_min_batch and B % self.lanes == 0) else 1

    def forward_cls(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float32, out: Optional[torch.Tensor] = None, *,
                    meta_tokens: Optional[torch.Tensor] = None, features: Optional[torch.Tensor] = None,
                    want_logits: bool = True, _ws=None) -> Optional[torch.Tensor]:
        """logits[B, num_classes] (``want_logits``) and / or the pre-head features[B, C_last] (bf16, written into ``features``).
        ``meta_tokens`` [B, M, C0]: the caller's own meta tokens (forward_features(x, c) of the reference); None = the model's."""
        self._check_open()
        x = self._prep_input(x)
        B, _, H, W = x.shape
        if meta_tokens is not None:
            require_cuda(meta_tokens)
            if tuple(meta_tokens.shape) != (B, self.cfg.queries_len, self.embed_dim[0]):
                raise RuntimeError(f"expected meta tokens [{B}, {self.cfg.queries_len}, {self.embed_dim[0]}], got {tuple(meta_tokens.shape)}")
            meta_tokens = meta_tokens.to(torch.bfloat16).contiguous()
        if features is not None and (features.dtype != torch.bfloat16 or tuple(features.shape) != (B, self.embed_dim[-1])
                                     or not features.is_contiguous()):
            raise RuntimeError("features buffer must be a contiguous bf16 [B, embed_dim[-1]] tensor")
        if want_logits and self.num_classes <= 0:
            raise RuntimeError("lemevit_b200: the model has no classifier (num_classes == 0); use forward_features")
        with torch.cuda.device(self.device):
            if want_logits and out is None:
                out = torch.empty((B, self.num_classes), dtype=out_dtype, device=self.device)
            lanes = self._use_lanes(B)
            cur = torch.cuda.current_stream(self.device)

            def launch(xi, ci, fi, oi, nb, ws, stream):
                _native.check(self.lib.lmv_forward_cls_features(
                    self._plan, xi.data_ptr(), self._x_code(x), nb, H, W, ci.data_ptr() if ci is not None else None,
                    ws.data_ptr(), ws.numel(), fi.data_ptr() if fi is not None else None,
                    oi.data_ptr() if oi is not None else None, _TORCH2LMV[oi.dtype] if oi is not None else _native.DTYPE_BF16,
                    stream.cuda_stream))

            if lanes == 1:
                ws = _ws[0] if _ws is not None else self._workspace(B, H, W)
                launch(x, meta_tokens, features, out if want_logits else None, B, ws, cur)
                return out if want_logits else None
            # independent sub-batches on concurrent streams: the images never interact, so the tail / latency
            # bubbles of one lane's kernels are filled by the other lane's CTAs.  Fork / join through events (graph-capturable).
            nb = B // lanes
            wss = _ws if _ws is not None else self._lane_workspaces(nb, H, W, lanes)
            fork = torch.cuda.Event()
            fork.record(cur)
            for i, st in enumerate(self._lane_streams(lanes)):
                st.wait_event(fork)
                sl = slice(i * nb, (i + 1) * nb)
                launch(x[sl], meta_tokens[sl] if meta_tokens is not None else None, features[sl] if features is not None else None,
                       out[sl] if want_logits else None, nb, wss[i], st)
                join = torch.cuda.Event()
                join.record(st)
                cur.wait_event(join)
        return out if want_logits else None

    def _lane_streams(self, lanes: int):
        if len(getattr(self, "_streams", [])) != lanes:
            self._streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]
        return self._streams

    def _lane_workspaces(self, nb: int, H: int, W: int, lanes: int):
        need = int(self.lib.lmv_workspace_bytes(self._plan, nb, H, W))
        if need == 0:
            raise RuntimeError(f"lemevit_b200: unsupported input shape {nb}x{H}x{W}")
        cur = getattr(self, "_lane_ws", [])
        if len(cur) != lanes or cur[0].numel() < need:
            self._lane_ws = [torch.empty(need, dtype=torch.uint8, device=self.device) for _ in range(lanes)]
        return self._lane_ws

    def forward_features(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float32,
                         outs: Optional[Sequence[torch.Tensor]] = None, _ws=None) -> List[torch.Tensor]:
        self._check_open()
        x = self._prep_input(x)
        B, _, H, W = x.shape
        with torch.cuda.device(self.device):
            ws = _ws[0] if _ws is not None else self._workspace(B, H, W)
            if outs is None:
                outs = [torch.empty(s, dtype=out_dtype, device=self.device) for s in self.out_shapes(B, H, W)]
            ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self.lib.lmv_forward_features(self._plan, x.data_ptr(), self._x_code(x), B, H, W,
                                                        ws.data_ptr(), ws.numel(), ptrs, len(outs),
                                                        _TORCH2LMV[outs[0].dtype], stream))
        return list(outs)

    # -- CUDA-graph replay of a fixed-shape forward ----------------------------------------------------
    def _graph_workspaces(self, B: int, H: int, W: int) -> List[torch.Tensor]:
        lanes = self._use_lanes(B)
        need = int(self.lib.lmv_workspace_bytes(self._plan, B // lanes, H, W))
        if need == 0:
            raise RuntimeError(f"lemevit_b200: unsupported input shape {B}x{H}x{W}")
        return [torch.empty(need, dtype=torch.uint8, device=self.device) for _ in range(lanes)]

    def graphed(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float32):
        """Capture the forward for x's shape once; returns (static_input, static_output(s), replay_fn).
        The graph owns its workspace; at most ``max_graphs`` shapes stay captured (least recently used evicted).  replay_fn
        raises once the engine was closed or the graph evicted — it never runs on memory that has been given back."""
        self._check_open()
        x = self._prep_input(x)
        key = (tuple(x.shape), x.dtype, self._x_code(x), out_dtype)
        if key not in self._graphs:
            B, _, H, W = x.shape
            static_x = x.clone()
            fwd = self.forward_features if self.backbone else self.forward_cls
            with torch.cuda.device(self.device):
                wss = self._graph_workspaces(B, H, W)
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    for _ in range(2):      # warm-up: builds the native schedule, sets kernel attributes
                        static_out = fwd(static_x, out_dtype, _ws=wss)
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    static_out = fwd(static_x, out_dtype, _ws=wss)
            while len(self._graphs) >= max(1, self.max_graphs):
                torch.cuda.synchronize(self.device)
                self._graphs.popitem(last=False)
            self._graphs[key] = (static_x, static_out, g, wss, self.packed)
        self._graphs.move_to_end(key)
        static_x, static_out, g, _, _ = self._graphs[key]
        gen = self._generation

        def replay():
            if self._closed or gen != self._generation or self._graphs.get(key, (None,) * 3)[2] is not g:
                raise RuntimeError("lemevit_b200: this CUDA graph is stale (engine closed, options changed or graph evicted); "
                                   "call graphed() again")
            g.replay()

        return static_x, static_out, replay
