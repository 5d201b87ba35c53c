"""Host-side mirror of the reference denoiser modules, backed by the CUDA library.

Same class names, constructor arguments, parameter names and call signatures as
``puzzle_diff/model/backbones/{efficient_gat,efficient_gat_3d,Transformer_GNN,
exophormer_gnn}.py`` so the classes drop into the reference's diffusion modules and
load reference checkpoints, but the ``nn.Parameter`` tensors here are only the
*storage* of the weights: every forward pass runs in ``libdiffassemble_b200.so``
through :class:`~diffassemble_b200.engine.DenoiserEngine`.  There is no PyTorch /
PyG execution path and no CPU path; calling a module with CPU tensors raises.

Out of scope: the visual / point-cloud encoders (attach your own
module as ``visual_backbone`` / ``pcd_backbone``), the ``gcn`` architecture, autograd
through the denoiser (training is row N1 of the scope table).
"""
import math
from typing import Optional

import torch
from torch import Tensor, nn

from . import _cabi
from .engine import DenoiserEngine


class TransformerConv(nn.Module):
    """Parameter holder with PyG ``TransformerConv`` key names (``lin_{key,query,value,skip}``)."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True):
        super().__init__()
        assert concat, "the reference only uses concat=True"
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        hc = heads * out_channels
        self.lin_key = nn.Linear(in_channels, hc)
        self.lin_query = nn.Linear(in_channels, hc)
        self.lin_value = nn.Linear(in_channels, hc)
        self.lin_skip = nn.Linear(in_channels, hc, bias=True)

    def forward(self, x, edge_index, return_attention_weights=None):
        """Single-layer execution through the CUDA operators (``da_op_linear`` +
        ``da_op_graph_attention``); the fused multi-layer path is ``Eff_GAT``."""
        from .engine import op_graph_attention, op_linear

        w = torch.cat([self.lin_query.weight, self.lin_key.weight, self.lin_value.weight, self.lin_skip.weight], 0)
        b = torch.cat([self.lin_query.bias, self.lin_key.bias, self.lin_value.bias, self.lin_skip.bias], 0)
        qkvs = op_linear(x, w.detach(), b.detach(), act=0, mode="fp32")
        if return_attention_weights:
            y, alpha = op_graph_attention(qkvs, edge_index, self.heads, return_alpha=True)
            return y, (edge_index, alpha)
        return op_graph_attention(qkvs, edge_index, self.heads)


def _conv_stack(input_size, hidden_dim, heads, output_size, n_layers):
    # Transformer_GNN.py:9-25 / exophormer_gnn.py:138-154
    return nn.ModuleList(
        [TransformerConv(input_size, out_channels=hidden_dim // heads, heads=heads)]
        + [TransformerConv(hidden_dim, out_channels=hidden_dim // heads, heads=heads) for _ in range(n_layers - 2)]
        + [TransformerConv(hidden_dim, heads=heads, concat=True, out_channels=output_size // heads)]
    )


class Transformer_GNN(nn.Module):
    """``Transformer_GNN.py:5-46``: parameter holder; executed fused inside ``Eff_GAT``."""

    arch = _cabi.DA_ARCH_TRANSFORMER

    def __init__(self, input_size, hidden_dim, heads, output_size, n_layers=4) -> None:
        super().__init__()
        self.module_list = _conv_stack(input_size, hidden_dim, heads, output_size, n_layers)
        self.n_layers = n_layers
        self.hidden_dim, self.heads = hidden_dim, heads
        self.virt_nodes = 0

    def extend_graph(self, edge_index: Tensor, batch: Tensor):
        return edge_index, len(batch), None


class Exophormer_GNN(nn.Module):
    """``exophormer_gnn.py:132-215``: parameter holder + the virtual-node wiring.

    ``extend_graph`` reproduces ``exophormer_gnn.py:164-200`` index for index (including
    the mis-aligned src/dst concatenation that wires real nodes of one graph to the
    virtual nodes of another when the batch holds several graphs), but is run once per
    batch instead of once per denoising step, and without the per-graph Python loop.
    """

    arch = _cabi.DA_ARCH_EXOPHORMER

    def __init__(self, input_size, hidden_dim, heads, output_size, n_layers=4, virt_nodes=4) -> None:
        super().__init__()
        self.module_list = _conv_stack(input_size, hidden_dim, heads, output_size, n_layers)
        self.virt_nodes = virt_nodes
        if self.virt_nodes > 0:
            self.virt_node_embedding = nn.Embedding(virt_nodes, input_size)
        self.n_layers = n_layers
        self.hidden_dim, self.heads = hidden_dim, heads

    def extend_graph(self, edge_index: Tensor, batch: Tensor):
        V = self.virt_nodes
        num_real = len(batch)
        if V <= 0:
            return edge_index, num_real, None
        dev = batch.device
        n_graphs = int(batch.max()) + 1
        # rows appended after the real ones carry embedding ids 0..V-1, repeated per graph (:169)
        virt_ids = torch.arange(V, device=dev).repeat(n_graphs)
        # per graph i the reference repeats its V virtual ids (n_i + V) times (:185-195, the
        # count is taken on the batch vector already extended with the virtual rows)
        counts = torch.bincount(batch, minlength=n_graphs) + V
        graph_of_rep = torch.repeat_interleave(torch.arange(n_graphs, device=dev), counts)
        base = num_real + graph_of_rep * V
        virt_edges = (base[:, None] + torch.arange(V, device=dev)[None, :]).reshape(-1)
        real = torch.arange(num_real, device=dev)
        src = torch.cat([real, virt_edges])  # :198
        dst = torch.cat([virt_edges, real])  # :199
        ext = torch.hstack((edge_index, torch.stack((src, dst))))  # :200
        return ext, num_real + V * n_graphs, virt_ids.to(torch.int32)


class _EngineMixin:
    """Caches one engine per device and re-syncs weights / graph / features only on change."""

    def _init_engine_state(self, gemm_mode, attn_mode):
        self.gemm_mode, self.attn_mode = gemm_mode, attn_mode
        self._engine: Optional[DenoiserEngine] = None
        self._weights_key = None
        self._graph_key = None
        self._feats_key = None
        # strong references to the tensors the graph / feature keys were taken from: while a key is cached its
        # tensors cannot be freed, so the caching allocator cannot hand their address (with _version 0 again) to
        # the NEXT batch's same-shaped tensors and make a different topology look like the cached one
        self._graph_refs = None
        # double buffering (prefetch): a second engine that a side stream fills with the NEXT batch
        self._spare: Optional[DenoiserEngine] = None
        self._spare_weights_key = None
        self._spare_last_use = None      # event on the compute stream after the spare engine's last launch
        self._prefetch_stream = None
        self._prefetched = None

    def _denoiser_state(self):
        skip = ("visual_backbone.", "pcd_backbone.", "linear1.", "linear2.")
        return {k: v for k, v in self.state_dict().items() if not k.startswith(skip) and k not in ("mean", "std")}

    def _tensor_key(self, t: Optional[Tensor]):
        if t is None:
            return None
        return (t.data_ptr(), tuple(t.shape), t._version, t.dtype, str(t.device))

    def _get_engine(self, device) -> DenoiserEngine:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} was called with tensors on {device}: the B200 denoiser has no CPU path"
            )
        dev = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        if self._engine is None or self._engine.device != dev:
            if self._engine is not None:
                self._engine.close()
            self._engine = self._make_engine(dev)
            self._weights_key = self._graph_key = self._feats_key = None
        wkey = tuple((k, v.data_ptr(), v._version) for k, v in self._denoiser_state().items())
        if wkey != self._weights_key:
            self._engine.load_weights(self._denoiser_state())
            self._weights_key = wkey
            self._feats_key = None
            if self.gnn_backbone.virt_nodes > 0:
                self._graph_key = None  # virtual rows hold embedding weights
        return self._engine

    def invalidate(self):
        """Force weights / graph / features to be re-sent on the next call."""
        self._weights_key = self._graph_key = self._feats_key = None
        self._graph_refs = None
        self._spare_weights_key = None
        self._prefetched = None

    def _bind(self, eng: DenoiserEngine, edge_index: Tensor, feats: Optional[Tensor], batch: Tensor):
        gkey = (self._tensor_key(edge_index), self._tensor_key(batch))
        if gkey != self._graph_key:
            ext, num_total, virt_ids = self.gnn_backbone.extend_graph(edge_index, batch)
            eng.set_graph(ext, batch, num_real=len(batch), num_total=num_total, virt_ids=virt_ids)
            self._ext_edge_index = ext  # what TransformerConv sees: the reference returns THIS with alpha
            self._graph_key = gkey
            self._graph_refs = (edge_index, batch)
            self._feats_key = None
        fkey = self._tensor_key(feats) if feats is not None else "zero"
        if fkey != self._feats_key:
            eng.set_features(feats)
            self._feats_key = fkey
        return eng

    def _weights_state_key(self):
        return tuple((k, v.data_ptr(), v._version) for k, v in self._denoiser_state().items())

    def prefetch(self, edge_index: Tensor, feats: Tensor, batch: Tensor, device=None):
        """Stage the NEXT batch while the current one computes (double buffering).

        The host -> device copies of ``edge_index`` / ``feats`` / ``batch`` (pinned host tensors copy
        asynchronously; ``edge_index`` may also be a deferred topology such as ``topology.ExpanderBatchSpec``, whose
        edge list is then written on the device and ``batch`` is ignored), the virtual-node wiring, ``da_set_graph`` (edge classification, bitmaps, residual CSR)
        and ``da_set_features`` (the step-invariant hoist GEMM) all run on a side stream into a second engine
        handle; none of it touches the stream or the handle the running sampling loop uses.  Returns the
        device tensors ``(edge_index, feats, batch)``: pass exactly these to ``p_sample_loop`` /
        ``forward_with_feats`` / ``p_sample`` next -- that call switches to the prepared engine after waiting
        (on the device, not the host) for the side stream.  The reference has no counterpart: it re-uploads and
        re-wires inside every step."""
        dev = torch.device(device if device is not None else (feats.device if feats.is_cuda else "cuda"))
        dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        main = torch.cuda.current_stream(dev)
        if self._prefetch_stream is None or self._prefetch_stream.device != dev:
            self._prefetch_stream = torch.cuda.Stream(dev)
        st = self._prefetch_stream
        if self._spare is None or self._spare.device != dev:
            if self._spare is not None:
                self._spare.close()
            self._spare = self._make_engine(dev)
            self._spare_weights_key, self._spare_last_use = None, None
        eng = self._spare
        with torch.cuda.stream(st):
            if self._spare_last_use is not None:
                st.wait_event(self._spare_last_use)   # its buffers may still be read by launches of the previous batch
            if any(isinstance(t_, Tensor) and t_.is_cuda for t_ in (edge_index, feats, batch)):
                # device inputs (e.g. the encoder's output for the next batch) may still be being written on the compute
                # stream: order this stream behind everything enqueued there so far.  Pinned host inputs keep full overlap.
                st.wait_stream(main)
            wkey = self._weights_state_key()
            if wkey != self._spare_weights_key:
                st.wait_stream(main)                  # parameters may have been written on the compute stream
                eng.load_weights(self._denoiser_state())
                self._spare_weights_key = wkey
            if hasattr(edge_index, "build"):
                # deferred topology (topology.ExpanderBatchSpec / DenseBatchSpec): written on the device, on this stream
                ei, b = edge_index.build(dev)
            else:
                ei = edge_index.to(dev, non_blocking=True)
                b = batch.to(dev, non_blocking=True)
            f = feats.to(dev, non_blocking=True)
            ext, num_total, virt_ids = self.gnn_backbone.extend_graph(ei, b)
            eng.set_graph(ext, b, num_real=len(b), num_total=num_total, virt_ids=virt_ids)
            eng.set_features(f)
            ready = torch.cuda.Event()
            ready.record(st)
        for t in (ei, b, f, ext):
            t.record_stream(main)   # allocated on the side stream, consumed on the compute stream
        self._prefetched = dict(gkey=(self._tensor_key(ei), self._tensor_key(b)), fkey=self._tensor_key(f), ready=ready,
                                ext=ext, wkey=wkey, refs=(ei, b, f))
        return ei, f, b

    def _take_prefetched(self, edge_index: Tensor, feats: Optional[Tensor], batch: Tensor) -> bool:
        pf = self._prefetched
        if pf is None or feats is None:
            return False
        if pf["gkey"] != (self._tensor_key(edge_index), self._tensor_key(batch)) or pf["fkey"] != self._tensor_key(feats):
            return False
        if pf["wkey"] != self._weights_state_key():
            self._prefetched = None   # weights changed since the batch was staged: fall back to the normal path
            return False
        main = torch.cuda.current_stream(edge_index.device)
        main.wait_event(pf["ready"])
        done = torch.cuda.Event()
        done.record(main)             # everything launched so far on the engine that now becomes the spare
        self._engine, self._spare = self._spare, self._engine
        self._weights_key, self._spare_weights_key = pf["wkey"], self._weights_key
        self._spare_last_use = done
        self._graph_key, self._feats_key = pf["gkey"], pf["fkey"]
        self._graph_refs = pf["refs"][:2]
        self._ext_edge_index = pf["ext"]
        self._prefetched = None
        return True

    def zero_feature_engine(self, edge_index: Tensor, batch: Tensor) -> DenoiserEngine:
        """A second engine bound to the same graph with all-zero features (the unconditional pass of classifier-free
        guidance, ``spatial_diffusion.py:578-586``), kept next to the main one so that neither re-binds per step."""
        dev = torch.device("cuda", edge_index.device.index if edge_index.device.index is not None else torch.cuda.current_device())
        eng = getattr(self, "_zero_engine", None)
        if eng is None or eng.device != dev:
            eng = self._make_engine(dev)
            self._zero_engine, self._zero_wkey, self._zero_gkey = eng, None, None
        wkey = self._weights_state_key()
        if wkey != self._zero_wkey:
            eng.load_weights(self._denoiser_state())
            self._zero_wkey, self._zero_gkey = wkey, None
        gkey = (self._tensor_key(edge_index), self._tensor_key(batch))
        if gkey != self._zero_gkey:
            ext, num_total, virt_ids = self.gnn_backbone.extend_graph(edge_index, batch)
            eng.set_graph(ext, batch, num_real=len(batch), num_total=num_total, virt_ids=virt_ids)
            eng.set_features(None)
            self._zero_gkey, self._zero_refs = gkey, (edge_index, batch)
        return eng

    def engine_for(self, edge_index: Tensor, feats: Optional[Tensor], batch: Tensor) -> DenoiserEngine:
        """Engine with this graph and these features bound (used by the fused sampler steps)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and feats is not None and feats.requires_grad:
            raise NotImplementedError("autograd through the CUDA denoiser is not implemented (scope row N1)")
        if self._prefetched is not None and self._take_prefetched(edge_index, feats, batch):
            return self._engine
        eng = self._get_engine(edge_index.device)
        return self._bind(eng, edge_index, feats, batch)


def _gnn(architecture, dim, n_layers, virt_nodes):
    if architecture == "transformer":
        return Transformer_GNN(dim, n_layers=n_layers, hidden_dim=32 * 8, heads=8, output_size=dim)
    if architecture == "exophormer":
        return Exophormer_GNN(dim, n_layers=n_layers, hidden_dim=32 * 8, heads=8, output_size=dim, virt_nodes=virt_nodes)
    raise NotImplementedError(f"architecture {architecture!r} is outside the B200 hot path (transformer | exophormer)")


class Eff_GAT(nn.Module, _EngineMixin):
    """``efficient_gat.py:15-146`` with the same constructor and call signatures."""

    def __init__(self, steps, input_channels=2, output_channels=2, n_layers=4, visual_pretrained=True,
                 freeze_backbone=False, model="efficientnet_b0", architecture="transformer", virt_nodes=4,
                 all_equivariant=False, gemm_mode="bf16x3", attn_mode="auto") -> None:
        super().__init__()
        # efficient_gat.py:40-42: timm's efficientnet_b0 feature pyramid -> the CUDA encoder of efficientnet.py (scope
        # row N4; random-init unless a checkpoint is loaded: there is no network for `visual_pretrained` weights).
        # Other encoders (resnet18 / resnet50 / resnet18equiv) stay upstream: attach a module or pass patch_feats.
        self.visual_backbone = None
        if model == "efficientnet_b0":
            from .efficientnet import EfficientNetB0Features

            self.visual_backbone = EfficientNetB0Features()
        self.all_equivariant = all_equivariant
        self.model = model
        self.combined_features_dim = {
            "resnet18": 3136, "resnet50": 12352, "efficientnet_b0": 1088 + 32 + 32, "resnet18equiv": 1088 + 32 + 32,
        }[model]
        self.input_channels, self.output_channels = input_channels, output_channels
        self.freeze_backbone = freeze_backbone
        self.steps = steps
        D = self.combined_features_dim
        self.gnn_backbone = _gnn(architecture, D, n_layers, virt_nodes)
        self.time_emb = nn.Embedding(steps, 32)
        self.pos_mlp = nn.Sequential(nn.Linear(input_channels, 16), nn.GELU(), nn.Linear(16, 32))
        self.final_mlp = nn.Sequential(nn.Linear(D, 32), nn.GELU(), nn.Linear(32, output_channels))
        self.mlp = nn.Sequential(nn.Linear(D, 128), nn.GELU(), nn.Linear(128, D))
        self.linear1 = nn.Linear(8192, 544)  # unused in the reference too; kept for checkpoint compatibility
        self.linear2 = nn.Linear(4096, 544)
        self.register_buffer("mean", torch.tensor([0.4850, 0.4560, 0.4060])[None, :, None, None])
        self.register_buffer("std", torch.tensor([0.2290, 0.2240, 0.2250])[None, :, None, None])
        self._init_engine_state(gemm_mode, attn_mode)

    def _make_engine(self, device):
        return DenoiserEngine(
            device=device, feat_dim=self.combined_features_dim - 64, in_channels=self.input_channels,
            out_channels=self.output_channels, steps=self.steps, mlp_hidden=128, head_kind=_cabi.DA_HEAD_2D,
            arch=self.gnn_backbone.arch, virt_nodes=self.gnn_backbone.virt_nodes, heads=self.gnn_backbone.heads,
            hidden=self.gnn_backbone.hidden_dim, n_layers=self.gnn_backbone.n_layers, gemm_mode=self.gemm_mode,
            attn_mode=self.attn_mode,
        )

    def forward(self, xy_pos, time, patch_rgb, edge_index, batch):
        patch_feats = self.visual_features(patch_rgb)
        return self.forward_with_feats(xy_pos, time, patch_rgb, edge_index, patch_feats=patch_feats, batch=batch)

    def forward_with_feats(self, xy_pos: Tensor, time: Tensor, patch_rgb: Tensor, edge_index: Tensor,
                           patch_feats: Tensor, batch, return_attention=False):
        if return_attention:
            return self._forward_with_attentions(xy_pos, time, edge_index, patch_feats, batch)
        eng = self.engine_for(edge_index, patch_feats, batch)
        return eng.forward(xy_pos, time), None

    def _forward_with_attentions(self, xy_pos, time, edge_index, patch_feats, batch):
        """``return_attention=True``: one ``(edge_index, alpha[E, H])`` tuple PER LAYER, as ``Transformer_GNN.forward``
        returns them (``Transformer_GNN.py:29-46``); ``Exophormer_GNN.forward`` only keeps the LAST layer's tuple, with
        the extended edge list that holds the virtual wiring (``exophormer_gnn.py:203-208``).  Attention weights per edge only exist on the CSR path, so this call
        runs on a second engine in ``attn_mode="csr"`` whatever the module's mode is (the tensor-core tiles never
        materialise per-edge weights); layer l's weights are obtained by truncating the stack after layer l -- the
        reference's viz / app code calls this once per sample, not per step."""
        eng = getattr(self, "_alpha_engine", None)
        dev = torch.device("cuda", edge_index.device.index if edge_index.device.index is not None else torch.cuda.current_device())
        if eng is None or eng.device != dev:
            saved = self.attn_mode
            self.attn_mode = "csr"
            try:
                eng = self._make_engine(dev)
            finally:
                self.attn_mode = saved
            self._alpha_engine, self._alpha_wkey = eng, None
        wkey = self._weights_state_key()
        if wkey != self._alpha_wkey:
            eng.load_weights(self._denoiser_state())
            self._alpha_wkey = wkey
        ext, num_total, virt_ids = self.gnn_backbone.extend_graph(edge_index, batch)
        eng.set_graph(ext, batch, num_real=len(batch), num_total=num_total, virt_ids=virt_ids)
        eng.set_features(patch_feats)
        out, alphas = eng.forward(xy_pos, time, return_alpha=True, all_layers=True)
        if self.gnn_backbone.arch == _cabi.DA_ARCH_EXOPHORMER:   # Exophormer_GNN.forward only keeps the last layer's (:203-208)
            return out, [(ext, alphas[-1])]
        return out, [(ext, a) for a in alphas]

    def visual_features(self, patch_rgb):
        if self.visual_backbone is None:
            raise NotImplementedError(
                "the CNN patch encoder is upstream of the B200 hot path (SURVEY.md section 2.1); "
                "attach a module as `visual_backbone` or pass pre-computed patch_feats"
            )
        from .efficientnet import EfficientNetB0Features

        if isinstance(self.visual_backbone, EfficientNetB0Features):
            # frozen backbone in eval mode, as the shipped configuration runs it (freeze_backbone=True, :153-155)
            self.visual_backbone.eval()
            feats = self.visual_backbone.forward(patch_rgb, self.mean, self.std)   # normalisation fused into the layout kernel
        else:
            feats = self.visual_backbone.forward((patch_rgb - self.mean) / self.std)
        return torch.cat([feats[2].reshape(patch_rgb.shape[0], -1), feats[3].reshape(patch_rgb.shape[0], -1)], -1)


class Eff_GAT_3d(nn.Module, _EngineMixin):
    """``efficient_gat_3d.py:48-237`` with the same constructor and call signatures."""

    FEAT_DIM = {"pointnet_inv": 1024, "pointnet": 128, "pointnet_plus": 256, "vn_dgcnn": 768, "vn_dgcnn_inv": 256, "vnn": 2104}

    def __init__(self, steps, input_channels=7, t_channels=3, r_channels=3, n_layers=4, architecture="transformer",
                 virt_nodes=8, backbone="pointnet", freeze_backbone=False, use_vn_dgcnn_equiv_inv_mp=False,
                 gemm_mode="bf16x3", attn_mode="auto") -> None:
        super().__init__()
        if use_vn_dgcnn_equiv_inv_mp:
            raise NotImplementedError("use_vn_dgcnn_equiv_inv_mp is outside the B200 hot path")
        if t_channels != 3 or r_channels != 3:
            raise NotImplementedError("only the 3 + 3 (translation, axis-angle) head is built")
        if backbone not in self.FEAT_DIM:
            raise Exception(f"Backbone not implemented {backbone}")
        feat_dim = self.FEAT_DIM[backbone]
        # efficient_gat_3d.py:77-79: the plain PointNet encoder runs on the CUDA operators (pointnet.py, scope row N4);
        # the other encoders (vn_dgcnn, pointnet_plus ...) stay upstream: attach one or pass pcd_feats
        self.pcd_backbone = None
        if backbone == "pointnet":
            from .pointnet import PointNet

            self.pcd_backbone = PointNet(feat_dim=feat_dim)
        self.combined_features_dim = feat_dim + 32 + 32
        self.gnn_feat_dim = self.combined_features_dim
        self.input_channels = input_channels
        self.freeze_backbone = freeze_backbone
        self.steps = steps
        D = self.gnn_feat_dim
        self.gnn_backbone = _gnn(architecture, D, n_layers, virt_nodes)
        self.time_emb = nn.Embedding(steps, 32)
        self.pos_mlp = nn.Sequential(nn.Linear(input_channels, 16), nn.GELU(), nn.Linear(16, 32))
        self.mlp = nn.Sequential(nn.Linear(D, 256), nn.LeakyReLU(0.2), nn.Linear(256, D), nn.LeakyReLU(0.2))
        self.mlp_t = nn.Sequential(nn.Linear(D, 256), nn.GELU(), nn.Linear(256, t_channels))
        self.mlp_r = nn.Sequential(nn.Linear(D, 256), nn.GELU(), nn.Linear(256, r_channels))
        self._init_engine_state(gemm_mode, attn_mode)

    def _make_engine(self, device):
        return DenoiserEngine(
            device=device, feat_dim=self.combined_features_dim - 64, in_channels=self.input_channels, out_channels=7,
            steps=self.steps, mlp_hidden=256, head_kind=_cabi.DA_HEAD_SE3, arch=self.gnn_backbone.arch,
            virt_nodes=self.gnn_backbone.virt_nodes, heads=self.gnn_backbone.heads, hidden=self.gnn_backbone.hidden_dim,
            n_layers=self.gnn_backbone.n_layers, gemm_mode=self.gemm_mode, attn_mode=self.attn_mode,
        )

    def forward(self, xy_pos, time, pcd, edge_index, batch):
        pcd_feats = self.pcd_features(pcd)
        return self.forward_with_feats(xy_pos, time, edge_index, pcd_feats=pcd_feats, batch=batch)

    def forward_with_feats(self, xy_pos: Tensor, time: Tensor, edge_index: Tensor, pcd_feats: Tensor, batch):
        eng = self.engine_for(edge_index, pcd_feats, batch)
        return eng.forward(xy_pos, time), None

    def pcd_features(self, pcd):
        if self.pcd_backbone is None:
            raise NotImplementedError(
                "the point-cloud encoder is upstream of the B200 hot path; attach `pcd_backbone` or pass pcd_feats"
            )
        return self.pcd_backbone(pcd)
