"""Training path of the denoiser (scope row N1): forward + backward through the CUDA operators.

Mirrors ``GNN_Diffusion.p_losses`` / ``training_step`` (``spatial_diffusion.py:432-483, 707-722``) and
``Eff_GAT.forward_with_feats`` (``efficient_gat.py:121-146``).  The heavy operators -- every linear layer
(forward, data gradient, weight gradient) and the TransformerConv attention stage (forward and backward over
CSR / CSC) -- are ``torch.autograd.Function`` s over the C-ABI kernels; torch autograd only chains them and
handles the small element-wise glue (GELU, concatenation, embedding lookup, Huber loss).  No PyG, no CPU path.

The optimizer is the reference's own choice, ``transformers.optimization.Adafactor()`` with default arguments
(``spatial_diffusion.py:701-705``), stepping the ``nn.Parameter`` storage that the inference engine also reads
(its packed copies are refreshed because the fused step bumps the parameters' version counters explicitly).
"""
import ctypes as C
import math
from typing import Optional

import torch
import torch.nn.functional as F

from . import _cabi
from ._cabi import DiffAssembleError
from .engine import _ptr, _require_cuda, _stream, op_linear


def _linear_mode(mode: str, N: int, K: int) -> str:
    # the tensor-core GEMM needs K % 64 == 0 and N % 32 == 0; tiny layers fall back to the exact-fp32 CUDA GEMM
    return mode if (mode == "bf16x3" and K % 64 == 0 and N % 32 == 0) else "fp32"


def op_linear_wgrad(dy: torch.Tensor, x: torch.Tensor, with_bias: bool = True):
    lib = _cabi.load_library()
    dy, x = dy.float().contiguous(), x.float().contiguous()
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
    db = torch.empty((N,), dtype=torch.float32, device=dy.device) if with_bias else None
    with torch.cuda.device(dy.device):
        st = lib.da_op_linear_wgrad(_ptr(dy), _ptr(x), _ptr(dw), _ptr(db), M, N, K, _stream(dy.device))
    if st != _cabi.DA_OK:
        raise DiffAssembleError(st, "da_op_linear_wgrad failed")
    return dw, db


class TrainGraph:
    """CSR-by-target + CSR-by-source of one batch's edge multiset (``da_graph``), built once per batch."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, batch: Optional[torch.Tensor] = None):
        _require_cuda(edge_index, "edge_index")
        self._lib = _cabi.load_library()
        ei = edge_index.to(torch.int64).contiguous()
        self.device, self.n, self.E = ei.device, int(num_nodes), ei.shape[1]
        h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            st = self._lib.da_graph_create(C.byref(h), _ptr(ei[0]), _ptr(ei[1]), self.E, self.n, _stream(self.device))
        if st != _cabi.DA_OK:
            raise DiffAssembleError(st, "da_graph_create failed (edge index out of range?)")
        self._h = h
        if batch is not None and batch.numel() == self.n:
            # dense-tile plan: lets the forward attention of dense puzzle graphs run on the tensor cores
            b = batch.to(device=self.device, dtype=torch.int64).contiguous()
            with torch.cuda.device(self.device):
                st = self._lib.da_graph_set_batch(self._h, _ptr(ei[0]), _ptr(ei[1]), _ptr(b), _stream(self.device))
            if st != _cabi.DA_OK:
                raise DiffAssembleError(st, "da_graph_set_batch failed")

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.da_graph_destroy(self._h)
                self._h = C.c_void_p(0)
        except Exception:
            pass

    def attention_fwd(self, qkvs: torch.Tensor, heads: int):
        qkvs = qkvs.float().contiguous()
        HC = qkvs.shape[1] // 4
        y = torch.empty((self.n, HC), dtype=torch.float32, device=self.device)
        stats = torch.empty((self.n, heads, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self._lib.da_op_graph_attention_fwd(self._h, _ptr(qkvs), heads, HC // heads, _ptr(y), _ptr(stats), _stream(self.device))
        if st != _cabi.DA_OK:
            raise DiffAssembleError(st, "da_op_graph_attention_fwd failed")
        return y, stats

    def attention_bwd(self, qkvs: torch.Tensor, stats: torch.Tensor, dy: torch.Tensor, heads: int):
        dy = dy.float().contiguous()
        HC = qkvs.shape[1] // 4
        dqkvs = torch.empty_like(qkvs)
        delta = torch.empty((self.n, heads), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self._lib.da_op_graph_attention_bwd(self._h, _ptr(qkvs), _ptr(stats), _ptr(dy), heads, HC // heads, _ptr(dqkvs),
                                                     _ptr(delta), _stream(self.device))
        if st != _cabi.DA_OK:
            raise DiffAssembleError(st, "da_op_graph_attention_bwd failed")
        return dqkvs


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, mode):
        x = x.float().contiguous()
        y = op_linear(x, w.detach(), b.detach() if b is not None else None, act=0, mode=_linear_mode(mode, w.shape[0], w.shape[1]))
        ctx.save_for_backward(x, w)
        ctx.mode, ctx.has_bias = mode, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.float().contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            wt = w.detach().t().contiguous()   # [K, N]: dx = dy @ w  ==  linear(dy, w^T)
            dx = op_linear(dy, wt, None, act=0, mode=_linear_mode(ctx.mode, wt.shape[0], wt.shape[1]))
        M, N, K = dy.shape[0], dy.shape[1], x.shape[1]
        if ctx.mode == "bf16x3" and M % 64 == 0 and K % 32 == 0 and N >= 32:
            # weight gradient on the tensor cores: dW[N, K] = dY^T X is the same linear operator with the node
            # dimension as its reduction axis (3-pass split-bf16, ~1e-5 relative); the two transposes are the only glue
            dw = op_linear(dy.t().contiguous(), x.t().contiguous(), None, act=0, mode="bf16x3")
            db = dy.sum(0) if ctx.has_bias else None
        else:
            dw, db = op_linear_wgrad(dy, x, with_bias=ctx.has_bias)
        return dx, dw, db, None


class _GraphAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkvs, graph: TrainGraph, heads: int):
        qkvs = qkvs.float().contiguous()
        y, stats = graph.attention_fwd(qkvs, heads)
        ctx.save_for_backward(qkvs, stats)
        ctx.graph, ctx.heads = graph, heads
        return y

    @staticmethod
    def backward(ctx, dy):
        qkvs, stats = ctx.saved_tensors
        return ctx.graph.attention_bwd(qkvs, stats, dy, ctx.heads), None, None


def linear(x, layer: torch.nn.Linear, mode: str):
    return _LinearFn.apply(x, layer.weight, layer.bias, mode)


def denoiser_forward_train(model, xy_pos, time, edge_index, patch_feats, batch, graph: Optional[TrainGraph] = None):
    """Differentiable ``Eff_GAT.forward_with_feats`` (2-D trunk): returns ``(out, graph)``.

    ``graph`` (the batch's ``TrainGraph``, including the exophormer virtual wiring) can be passed back in on the
    next call with the same topology to skip the CSR / CSC build."""
    mode = model.gemm_mode
    gnn = model.gnn_backbone
    num_real = len(batch)
    if graph is None:
        ext, num_total, virt_ids = gnn.extend_graph(edge_index, batch)
        # (the dense-tile forward needs every node in a graph's tile: batches without virtual rows)
        graph = TrainGraph(ext, num_total, batch if (num_total == num_real and model.attn_mode == "auto") else None)
        graph.virt_ids = virt_ids
    time_feats = F.embedding(time, model.time_emb.weight)                               # efficient_gat.py:131
    pos_feats = linear(F.gelu(linear(xy_pos, model.pos_mlp[0], mode)), model.pos_mlp[2], mode)   # :132
    combined = torch.cat([patch_feats, pos_feats, time_feats], -1)                       # :134
    combined = linear(F.gelu(linear(combined, model.mlp[0], mode)), model.mlp[2], mode)  # :135
    x = combined
    if getattr(graph, "virt_ids", None) is not None:                                     # exophormer_gnn.py:169-178
        x = torch.cat((x, F.embedding(graph.virt_ids.long(), gnn.virt_node_embedding.weight)))
    n_layers = gnn.n_layers
    for l, conv in enumerate(gnn.module_list):
        w = torch.cat([conv.lin_query.weight, conv.lin_key.weight, conv.lin_value.weight, conv.lin_skip.weight], 0)
        b = torch.cat([conv.lin_query.bias, conv.lin_key.bias, conv.lin_value.bias, conv.lin_skip.bias], 0)
        qkvs = _LinearFn.apply(x, w, b, mode)
        x = _GraphAttentionFn.apply(qkvs, graph, conv.heads)
        if l < n_layers - 1 and gnn.arch == _cabi.DA_ARCH_TRANSFORMER:
            x = F.gelu(x)                                                                # Transformer_GNN.py:36
    feats = x[:num_real]
    out = linear(F.gelu(linear(feats + combined, model.final_mlp[0], mode)), model.final_mlp[2], mode)   # :144
    return out, graph


def allreduce_gradients(parameters, world_size: int, group=None):
    """DDP-style gradient averaging for the sharded training step: one flat NCCL all-reduce."""
    import torch.distributed as dist

    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads or world_size == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= world_size
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


# ---------------------------------------------------------------------------------------------
# Fused Adafactor
# ---------------------------------------------------------------------------------------------
def _adafactor_base():
    from transformers.optimization import Adafactor

    return Adafactor


def FusedAdafactor(params, **kwargs):
    """``transformers.optimization.Adafactor`` (the reference's optimizer, ``spatial_diffusion.py:701-705``) with its
    ``step`` replaced by ONE launch of ``da_adafactor_step`` for every fp32 CUDA parameter of rank <= 2 (all of the
    denoiser's).  Same constructor arguments, same state entries (``step``, ``exp_avg_sq_row``, ``exp_avg_sq_col``,
    ``exp_avg_sq``, ``RMS``), so optimizer checkpoints are interchangeable.  Parameters the kernel does not cover
    (rank > 2, first-moment runs, CPU) take the stock code path inside the same ``step`` call."""
    import ctypes as C
    import math

    import numpy as np

    from . import _cabi

    Base = _adafactor_base()

    class _FusedAdafactor(Base):
        _DESC = np.dtype([("p", "<u8"), ("g", "<u8"), ("sq_row", "<u8"), ("sq_col", "<u8"), ("sq", "<u8"), ("rms", "<u8"),
                          ("rows", "<i4"), ("cols", "<i4"), ("beta2t", "<f4"), ("rel_step", "<f4")])

        @torch.no_grad()
        def step(self, closure=None):
            loss = closure() if closure is not None else None
            lib = _cabi.load_library()
            stash = []
            for group in self.param_groups:
                fused = []
                for p in group["params"]:
                    g = p.grad
                    ok = (g is not None and p.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32 and p.dim() <= 2
                          and p.dim() >= 1 and group["beta1"] is None and group["relative_step"] and group["scale_parameter"]
                          and not group["warmup_init"] and p.is_contiguous() and not g.is_sparse)
                    if ok:
                        fused.append(p)
                if not fused:
                    continue
                dev = fused[0].device
                desc = np.zeros(len(fused), dtype=self._DESC)
                for k, p in enumerate(fused):
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = 0
                        if p.dim() == 2:
                            st["exp_avg_sq_row"] = torch.zeros(p.shape[0], dtype=torch.float32, device=dev)
                            st["exp_avg_sq_col"] = torch.zeros(p.shape[1], dtype=torch.float32, device=dev)
                        else:
                            st["exp_avg_sq"] = torch.zeros_like(p)
                        st["RMS"] = torch.zeros((), dtype=torch.float32, device=dev)
                    if not torch.is_tensor(st["RMS"]):
                        st["RMS"] = torch.zeros((), dtype=torch.float32, device=dev)
                    st["step"] += 1
                    d = desc[k]
                    d["p"], d["g"], d["rms"] = p.data_ptr(), g.data_ptr(), st["RMS"].data_ptr()
                    if p.dim() == 2:
                        d["sq_row"], d["sq_col"] = st["exp_avg_sq_row"].data_ptr(), st["exp_avg_sq_col"].data_ptr()
                        d["rows"], d["cols"] = p.shape
                    else:
                        d["sq"] = st["exp_avg_sq"].data_ptr()
                        d["rows"], d["cols"] = 1, p.numel()
                    d["beta2t"] = 1.0 - math.pow(st["step"], group["decay_rate"])
                    d["rel_step"] = min(1e-2, 1.0 / math.sqrt(st["step"]))
                    stash.append((p, p.grad, g))
                table = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).to(dev, non_blocking=True)
                with torch.cuda.device(dev):
                    stt = lib.da_adafactor_step(C.c_void_p(table.data_ptr()), len(fused), float(group["eps"][0]),
                                                float(group["eps"][1]), float(group["clip_threshold"]),
                                                float(group["weight_decay"]), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
                if stt != _cabi.DA_OK:
                    raise _cabi.DiffAssembleError(stt, "da_adafactor_step failed")
                table.record_stream(torch.cuda.current_stream(dev))
                # the kernel wrote the parameters through raw pointers: bump their version counters so that every
                # cache keyed on (data_ptr, _version) -- the inference engine's packed weight copies, the prefetch
                # engine, captured loop graphs -- sees the update (an in-place torch op would have done this itself)
                torch.autograd.graph.increment_version(fused)
            # everything else: the stock implementation, with the fused parameters' gradients hidden from it
            for p, grad, _ in stash:
                p.grad = None
            try:
                if any(q.grad is not None for group in self.param_groups for q in group["params"]):
                    super().step()
            finally:
                for p, grad, _ in stash:
                    p.grad = grad
            return loss

    return _FusedAdafactor(params, **kwargs)
