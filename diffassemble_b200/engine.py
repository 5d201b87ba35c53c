"""Thin PyTorch shim over the C-ABI handle: owns nothing but pointers.

PyTorch is plumbing here (device memory, streams); all arithmetic of the denoiser
runs in ``libdiffassemble_b200.so``.  Tensors passed in must be CUDA, contiguous
and of the exact dtype the ABI names; the shim converts dtype / contiguity only.
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import _cabi
from ._cabi import DiffAssembleError, da_config, da_step_coef, da_weight_desc


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} is on {t.device}: the B200 denoiser has no CPU path (move the module and its inputs to cuda)"
        )


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DenoiserEngine:
    """One ``da_handle``: weights + one graph + one feature set, on one device."""

    def __init__(self, *, device, feat_dim, in_channels, out_channels, steps, mlp_hidden, head_kind, arch,
                 virt_nodes=0, heads=8, hidden=256, n_layers=4, gemm_mode="bf16x3", attn_mode="auto"):
        self._lib = _cabi.load_library()
        self._h = C.c_void_p(0)
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DenoiserEngine needs a cuda device: there is no CPU fallback")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        cfg = da_config()
        cfg.abi_version = _cabi.DA_ABI_VERSION
        cfg.device = self.device.index
        cfg.feat_dim, cfg.in_channels, cfg.out_channels = feat_dim, in_channels, out_channels
        cfg.heads, cfg.hidden, cfg.n_layers, cfg.steps = heads, hidden, n_layers, steps
        cfg.mlp_hidden, cfg.head_kind, cfg.arch, cfg.virt_nodes = mlp_hidden, head_kind, arch, virt_nodes
        cfg.gemm_mode = _cabi.GEMM_MODES[gemm_mode] if isinstance(gemm_mode, str) else int(gemm_mode)
        cfg.attn_mode = _cabi.ATTN_MODES[attn_mode] if isinstance(attn_mode, str) else int(attn_mode)
        self.cfg = cfg
        self.out_channels, self.in_channels, self.heads = out_channels, in_channels, heads
        h = C.c_void_p(0)
        st = self._lib.da_create(C.byref(h), C.byref(cfg))
        if st != _cabi.DA_OK:
            raise DiffAssembleError(st, self._lib.da_last_error(None).decode())
        self._h = h
        self.num_real = 0
        self.num_total = 0
        self.num_edges = 0
        self._keep = {}  # tensors the library reads asynchronously

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.da_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _cabi.DA_OK:
            raise DiffAssembleError(st, self._lib.da_last_error(self._h).decode())

    # -- state -------------------------------------------------------------------------------
    def load_weights(self, state: Dict[str, torch.Tensor]):
        """``state``: denoiser state_dict entries (reference key names, prefix stripped)."""
        descs, keep = [], []
        for name, t in state.items():
            if not torch.is_floating_point(t):
                continue
            t = t.detach().to(dtype=torch.float32).contiguous()
            keep.append(t)
            if t.dim() == 1:
                rows, cols = t.shape[0], 1
            elif t.dim() == 2:
                rows, cols = t.shape
            else:
                continue
            descs.append(da_weight_desc(name.encode(), C.c_void_p(t.data_ptr()), rows, cols))
        arr = (da_weight_desc * len(descs))(*descs)
        with torch.cuda.device(self.device):
            self._check(self._lib.da_load_weights(self._h, arr, len(descs)))

    def set_graph(self, edge_index: torch.Tensor, batch: torch.Tensor, num_real: int, num_total: Optional[int] = None,
                  virt_ids: Optional[torch.Tensor] = None):
        _require_cuda(edge_index, "edge_index")
        ei = edge_index.to(dtype=torch.int64).contiguous()
        b = batch.to(device=self.device, dtype=torch.int64).contiguous()
        num_total = num_real if num_total is None else num_total
        vi = None
        if virt_ids is not None:
            vi = virt_ids.to(device=self.device, dtype=torch.int32).contiguous()
        E = ei.shape[1]
        src, dst = ei[0], ei[1]
        self._check(self._lib.da_set_graph(self._h, _ptr(src), _ptr(dst), E, _ptr(b), num_real, num_total, _ptr(vi),
                                           _stream(self.device)))
        self.num_real, self.num_total, self.num_edges = num_real, num_total, E

    def set_features(self, feats: Optional[torch.Tensor]):
        if feats is not None:
            _require_cuda(feats, "node features")
            feats = feats.to(dtype=torch.float32).contiguous()
            if feats.shape != (self.num_real, self.cfg.feat_dim):
                raise ValueError(f"features have shape {tuple(feats.shape)}, expected {(self.num_real, self.cfg.feat_dim)}")
        self._keep["feats"] = feats
        self._check(self._lib.da_set_features(self._h, _ptr(feats), _stream(self.device)))

    # -- compute -----------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, t: torch.Tensor, return_alpha: bool = False, all_layers: bool = False):
        _require_cuda(x, "x")
        x = x.to(dtype=torch.float32).contiguous()
        t = t.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty((self.num_real, self.out_channels), dtype=torch.float32, device=self.device)
        if return_alpha and all_layers:   # one [E, H] tensor per layer (da_forward_attn)
            alphas = torch.empty((self.cfg.n_layers, self.num_edges, self.heads), dtype=torch.float32, device=self.device)
            self._check(self._lib.da_forward_attn(self._h, _ptr(x), _ptr(t), _ptr(out), _ptr(alphas), _stream(self.device)))
            return out, list(alphas.unbind(0))
        alpha = None
        if return_alpha:
            alpha = torch.empty((self.num_edges, self.heads), dtype=torch.float32, device=self.device)
        self._check(self._lib.da_forward(self._h, _ptr(x), _ptr(t), _ptr(out), _ptr(alpha), _stream(self.device)))
        return (out, alpha) if return_alpha else out

    def _step(self, fn, x, coef: da_step_coef, noise, out):
        _require_cuda(x, "x")
        x = x.to(dtype=torch.float32).contiguous()
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        if out is None:
            out = torch.empty_like(x)
        self._check(fn(self._h, _ptr(x), _ptr(out), C.byref(coef), _ptr(noise), _stream(self.device)))
        return out

    def ddpm_step(self, x, coef, noise=None, out=None):
        return self._step(self._lib.da_ddpm_step, x, coef, noise, out)

    def ddim_step(self, x, coef, noise=None, out=None):
        return self._step(self._lib.da_ddim_step, x, coef, noise, out)

    def _step_t(self, x, t, noise, out):
        _require_cuda(x, "x")
        x = x.to(dtype=torch.float32).contiguous()
        t = t.to(device=self.device, dtype=torch.int64).contiguous()
        if t.shape[0] != self.num_real:
            raise ValueError(f"t has {t.shape[0]} entries, expected one per node ({self.num_real})")
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        return x, t, noise, (torch.empty_like(x) if out is None else out)

    def ddpm_step_t(self, x, t, t_index: int, sched, noise=None, out=None):
        """``da_ddpm_step_t``: per-node timesteps, schedule coefficients gathered on the device (no host look at t)."""
        x, t, noise, out = self._step_t(x, t, noise, out)
        self._keep["sched"] = sched
        self._check(self._lib.da_ddpm_step_t(self._h, _ptr(x), _ptr(out), _ptr(t), int(t_index), C.byref(sched.struct), _ptr(noise),
                                             _stream(self.device)))
        return out

    def ddim_step_t(self, x, t, pred: int, eta: float, sched, noise=None, out=None):
        x, t, noise, out = self._step_t(x, t, noise, out)
        self._keep["sched"] = sched
        self._check(self._lib.da_ddim_step_t(self._h, _ptr(x), _ptr(out), _ptr(t), int(pred), float(eta), C.byref(sched.struct), _ptr(noise),
                                             _stream(self.device)))
        return out

    def ddim_update(self, x, model_out, coef, noise=None):
        out = torch.empty_like(x)
        self._check(self._lib.da_ddim_update(self._h, _ptr(x.contiguous()), _ptr(model_out.contiguous()), _ptr(out),
                                             C.byref(coef), _ptr(noise), _stream(self.device)))
        return out

    # -- introspection -----------------------------------------------------------------------
    def workspace_bytes(self) -> int:
        return int(self._lib.da_workspace_bytes(self._h))

    def launch_count(self) -> int:
        return int(self._lib.da_launch_count(self._h))

    def graph_stats(self):
        nd, nc, ng = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        self._check(self._lib.da_graph_stats(self._h, C.byref(nd), C.byref(nc), C.byref(ng)))
        return {"dense_edges": nd.value, "csr_edges": nc.value, "dense_graphs": ng.value}

    def plan_info(self):
        """What the dense-tile planner made of the bound graph (``da_graph_plan_info``)."""
        out = (C.c_int64 * 11)()
        self._check(self._lib.da_graph_plan_info(self._h, out, 11))
        keys = ("tiles", "blocks_total", "blocks_visited", "blocks_full", "reordered_graphs", "extra_sources", "fused_rows", "csr_rows",
                "real_rows_clean", "folded", "persistent_hidden_launches")
        return dict(zip(keys, [int(v) for v in out]))

    def set_profiling(self, enable: bool):
        self._check(self._lib.da_set_profiling(self._h, 1 if enable else 0))

    def get_profile(self, reset: bool = True):
        n = 32
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        ntags = self._lib.da_get_profile(self._h, ms, cnt, n, 1 if reset else 0)
        return {
            self._lib.da_profile_tag_name(i).decode(): {"ms": ms[i], "launches": cnt[i]}
            for i in range(min(n, ntags))
        }


# -- stand-alone operators (unit-level parity tests) ----------------------------------------------
def op_linear(a, w, bias=None, act=0, mode="fp32"):
    """y = act(a @ w^T + bias) on the device; scratch for the split-bf16 planes comes from torch's caching
    allocator (stream-ordered), so the call neither allocates with cudaMalloc nor synchronises."""
    lib = _cabi.load_library()
    _require_cuda(a, "a")
    a = a.float().contiguous()
    w = w.float().contiguous()
    b = bias.float().contiguous() if bias is not None else None
    M, K = a.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=a.device)
    m = _cabi.GEMM_MODES[mode]
    nbytes = int(lib.da_op_linear_workspace_bytes(m, M, N, K))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device) if nbytes else None
    with torch.cuda.device(a.device):
        st = lib.da_op_linear_ws(m, _ptr(a), _ptr(w), _ptr(b), _ptr(y), M, N, K, act, _ptr(ws), nbytes, _stream(a.device))
    if st != _cabi.DA_OK:
        raise DiffAssembleError(st, "da_op_linear_ws failed")
    return y


def op_segment_max(x, seg_ptr):
    """``out[g] = x[seg_ptr[g]:seg_ptr[g + 1]].max(0)`` on the device (PointNet's global max pool)."""
    lib = _cabi.load_library()
    _require_cuda(x, "x")
    x = x.float().contiguous()
    sp = seg_ptr.to(device=x.device, dtype=torch.int32).contiguous()
    n_seg = sp.numel() - 1
    out = torch.empty((n_seg, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = lib.da_op_segment_max(_ptr(x), x.shape[1], _ptr(sp), n_seg, x.shape[1], _ptr(out), _stream(x.device))
    if st != _cabi.DA_OK:
        raise DiffAssembleError(st, "da_op_segment_max failed")
    return out


def op_graph_attention(qkvs, edge_index, heads, return_alpha=False):
    lib = _cabi.load_library()
    _require_cuda(qkvs, "qkvs")
    qkvs = qkvs.float().contiguous()
    n = qkvs.shape[0]
    HC = qkvs.shape[1] // 4
    ei = edge_index.to(device=qkvs.device, dtype=torch.int64).contiguous()
    E = ei.shape[1]
    y = torch.empty((n, HC), dtype=torch.float32, device=qkvs.device)
    alpha = torch.empty((E, heads), dtype=torch.float32, device=qkvs.device) if return_alpha else None
    with torch.cuda.device(qkvs.device):
        st = lib.da_op_graph_attention(_ptr(qkvs), _ptr(ei[0]), _ptr(ei[1]), E, n, heads, HC // heads, _ptr(y),
                                       _ptr(alpha), _stream(qkvs.device))
    if st != _cabi.DA_OK:
        raise DiffAssembleError(st, "da_op_graph_attention failed")
    return (y, alpha) if return_alpha else y


def op_graph_attention_dense(qkvs, edge_index, batch, heads):
    """Tensor-core dense-tile path + residual CSR; returns ``(y, n_dense_edges)``."""
    lib = _cabi.load_library()
    _require_cuda(qkvs, "qkvs")
    qkvs = qkvs.float().contiguous()
    n = qkvs.shape[0]
    HC = qkvs.shape[1] // 4
    ei = edge_index.to(device=qkvs.device, dtype=torch.int64).contiguous()
    b = batch.to(device=qkvs.device, dtype=torch.int64).contiguous()
    y = torch.empty((n, HC), dtype=torch.float32, device=qkvs.device)
    nd = C.c_int64(0)
    with torch.cuda.device(qkvs.device):
        st = lib.da_op_graph_attention_dense(_ptr(qkvs), _ptr(ei[0]), _ptr(ei[1]), ei.shape[1], _ptr(b), n, heads,
                                             HC // heads, _ptr(y), C.byref(nd), _stream(qkvs.device))
    if st != _cabi.DA_OK:
        raise DiffAssembleError(st, "da_op_graph_attention_dense failed")
    return y, nd.value
