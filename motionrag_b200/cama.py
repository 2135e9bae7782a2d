"""`CamaTransformer` — the CAMA causal motion transformer forward on libmrag kernels.

Takes the weights of the reference's `torch.nn.TransformerEncoder`
(configs/cogvideox/MotionRAG_open.yml:253-267: 4 post-norm layers, d_model 1024, 16 heads,
dim_feedforward 4096, gelu, batch_first) and runs `transformer(x, get_mask(G, L))`
(src/projects/condition/module.py:303-306) as 7 kernels per layer — tcgen05 GEMMs (K5), a
block-causal attention (K6) and fused residual + LayerNorm (K7) — replayed as one CUDA graph.
The input buffer is owned by the handle so that `gather_context(..., out=cama.input_view(b))` can
build the context in place. There is no PyTorch fallback: without libmrag this raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import CamaLayer, check
from .store import _stream_ptr, _view


class CamaTransformer:
    def __init__(self, encoder, groups: int = 10, group_tokens: int = 25, max_batch: int = 16,
                 device: int | str | torch.device = 0):
        """encoder: a torch.nn.TransformerEncoder (any device / dtype; weights are copied as bf16)."""
        self._lib = _cabi.load()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise _cabi.MragError(-3, "CamaTransformer runs on a CUDA device only")
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        layers = list(encoder.layers)
        first = layers[0]
        if getattr(encoder, "norm", None) is not None:
            raise ValueError("a final encoder norm is not part of the reference configuration")
        if first.norm_first or not first.self_attn.batch_first or first.activation_relu_or_gelu != 2:
            raise ValueError("expected post-norm, batch_first, gelu encoder layers (the reference's CAMA config)")
        if abs(first.norm1.eps - 1e-5) > 1e-12:
            raise ValueError("LayerNorm eps must be 1e-5")
        self.d_model = first.self_attn.embed_dim
        self.n_heads = first.self_attn.num_heads
        self.d_ff = first.linear1.out_features
        self.groups, self.group_tokens, self.max_batch = int(groups), int(group_tokens), int(max_batch)
        self.tokens = self.groups * self.group_tokens
        self._w = []  # keeps the bf16 copies alive

        def w(t):
            t = t.detach().to(self.device, torch.bfloat16).contiguous()
            self._w.append(t)
            return t.data_ptr()

        arr = (CamaLayer * len(layers))()
        for i, l in enumerate(layers):
            arr[i] = CamaLayer(w(l.self_attn.in_proj_weight), w(l.self_attn.in_proj_bias),
                               w(l.self_attn.out_proj.weight), w(l.self_attn.out_proj.bias),
                               w(l.linear1.weight), w(l.linear1.bias), w(l.linear2.weight), w(l.linear2.bias),
                               w(l.norm1.weight), w(l.norm1.bias), w(l.norm2.weight), w(l.norm2.bias))
        self._layers = arr
        h = C.c_void_p()
        check(self._lib.mrag_cama_create(len(layers), arr, self.d_model, self.n_heads, self.d_ff, self.groups,
                                         self.group_tokens, self.max_batch, self.device.index, C.byref(h)))
        self._h = h
        xin, yout = C.c_void_p(), C.c_void_p()
        check(self._lib.mrag_cama_io(self._h, C.byref(xin), C.byref(yout)))
        shape = (self.max_batch, self.tokens, self.d_model)
        self.x_in = _view(xin.value, shape, torch.bfloat16, self.device, self)
        self.y_out = _view(yout.value, shape, torch.bfloat16, self.device, self)
        self._y_last = None

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mrag_cama_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def input_view(self, b: int) -> torch.Tensor:
        """[b, tokens, d_model] bf16 view of the input buffer (target for gather_context(out=...))."""
        return self.x_in[:b]

    def forward(self, x: torch.Tensor | None = None, b: int | None = None, use_graph: bool = True) -> torch.Tensor:
        """transformer(x, block_causal_mask) -> [b, tokens, d_model] bf16 (a view of the output buffer,
        valid until the next call). Pass x=None when the input buffer was filled in place."""
        if x is not None:
            b = x.shape[0]
            if tuple(x.shape[1:]) != (self.tokens, self.d_model):
                raise ValueError(f"x must be [b, {self.tokens}, {self.d_model}]")
            self.x_in[:b].copy_(x)
        if b is None:
            raise ValueError("pass x or b")
        check(self._lib.mrag_cama_forward(self._h, int(b), 1 if use_graph else 0, _stream_ptr(self.device)))
        return self.y_out[:b]

    def predict(self, x: torch.Tensor | None = None, b: int | None = None, use_graph: bool = True) -> torch.Tensor:
        """ActionTransformer.predict's slice (module.py:326): the last group's tokens [b, L, d] — a view of the
        handle's prediction buffer, valid until the next call. The last layer only computes that group's rows
        (mrag_cama_predict); the values equal forward(...)[:, -L:] up to the fp32 summation order of the split-K GEMMs."""
        if x is not None:
            b = x.shape[0]
            if tuple(x.shape[1:]) != (self.tokens, self.d_model):
                raise ValueError(f"x must be [b, {self.tokens}, {self.d_model}]")
            self.x_in[:b].copy_(x)
        if b is None:
            raise ValueError("pass x or b")
        ptr = C.c_void_p()
        check(self._lib.mrag_cama_predict(self._h, int(b), 1 if use_graph else 0, _stream_ptr(self.device), C.byref(ptr)))
        if self._y_last is None or self._y_last.data_ptr() != ptr.value:
            self._y_last = _view(ptr.value, (self.max_batch, self.group_tokens, self.d_model), torch.bfloat16, self.device, self)
        return self._y_last[:b]


def linear(a: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, gelu: bool = False,
           splits: int = 1, cluster_reduce: bool = False, partial: bool = False) -> torch.Tensor:
    """The K5 GEMM on its own: a [M,K] bf16, weight [N,K] bf16 -> bf16 [M,N] (splits == 1, or
    cluster_reduce: the K splits are summed on chip inside a thread-block cluster) or fp32 partial sums
    [splits, M, N] (always when partial=True)."""
    lib = _cabi.load()
    M, K = a.shape
    N = weight.shape[0]
    dev = a.device
    a = a.contiguous()
    rows_alloc = M
    if M < 128:   # the TMA box is 128 rows tall: give it a buffer at least that tall
        pad = torch.zeros((128, K), dtype=a.dtype, device=dev)
        pad[:M] = a
        a, rows_alloc = pad, 128
    if (splits == 1 and not partial) or cluster_reduce:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
        part = None
    else:
        out = None
        part = torch.empty((splits, M, N), dtype=torch.float32, device=dev)
    check(lib.mrag_linear(C.c_void_p(a.data_ptr()), rows_alloc, C.c_void_p(weight.data_ptr()), M, N, K,
                          C.c_void_p(bias.data_ptr()) if bias is not None else None, 1 if gelu else 0,
                          C.c_void_p(out.data_ptr()) if out is not None else None,
                          C.c_void_p(part.data_ptr()) if part is not None else None, splits, _stream_ptr(dev)))
    return out if out is not None else part
