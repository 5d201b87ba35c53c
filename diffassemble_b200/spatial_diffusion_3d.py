"""Host-side mirror of ``puzzle_diff/model/spatial_diffusion_3d_test_double_diffusion.py``
(the ``GNN_Diffusion`` that ``train_3d.py:19`` uses): R^3 diffusion on the translation +
SO(3) diffusion on the rotation, DDIM only (``:275-281``).  Each sampling step is one fused
``da_ddim_step`` call; the SO(3) log/exp maps run in closed form on the device
(``csrc/se3.cuh``) instead of ``torch.matrix_exp`` / ``torch.linalg.eigh`` chains.
"""
from functools import partial
from typing import Any, Optional

import torch
from torch import Tensor, nn

from . import _cabi
from .backbones import Eff_GAT_3d
from .spatial_diffusion import DiffusionScheduleMixin, ModelMeanType, ModelScheduler, _Base, _HAVE_PL


class GNN_Diffusion_3d(_Base, DiffusionScheduleMixin):
    def __init__(
        self,
        steps=600,
        inference_ratio=1,
        sampling="DDPM",
        learning_rate=1e-4,
        save_and_sample_every=1000,
        classifier_free_prob=0,
        classifier_free_w=0,
        noise_weight=0.0,
        model_mean_type: ModelMeanType = ModelMeanType.EPSILON,
        input_channels=7,
        output_channels=7,
        scheduler: ModelScheduler = ModelScheduler.LINEAR,
        visual_pretrained: bool = True,
        freeze_backbone: bool = True,
        n_layers: int = 4,
        loss_type="all",
        backbone="vnn",
        max_epochs=200,
        use_vn_dgcnn_equiv_inv_mp: bool = False,
        max_num_part: int = 20,
        use_6dof: bool = False,
        architecture="transformer",
        gemm_mode: str = "bf16x3",
        attn_mode: str = "auto",
        *args,
        **kwargs,
    ) -> None:
        super().__init__(*args, **kwargs)
        if use_6dof:
            raise NotImplementedError("use_6dof is outside the B200 hot path")
        self.loss_type = loss_type
        self.free_backbone = freeze_backbone
        self.model_mean_type = model_mean_type
        self.learning_rate = learning_rate
        self.noise_weight = noise_weight
        self.backbone = backbone
        self.max_num_part = max_num_part
        self.use_6dof = use_6dof
        self.save_eval_images = False
        self.gemm_mode, self.attn_mode = gemm_mode, attn_mode
        self.inference_ratio = inference_ratio
        self.sampling = sampling
        if sampling == "DDIM":  # the only sampler the reference binds (:275-281)
            self.p_sample = partial(self._p_sample, sampling_func=self.p_sample_ddim)
        self.eta = 0
        self._register_schedule(steps, scheduler)
        self.register_buffer("identity", torch.eye(3))
        self.steps = steps
        self.input_channels = input_channels
        self.architecture = architecture
        self.n_layers = n_layers
        self.init_backbone()
        if _HAVE_PL:
            self.save_hyperparameters()

    if not _HAVE_PL:

        @property
        def device(self):
            return self.betas.device

        local_rank = 0

    def init_backbone(self):  # :334-345
        self.model = Eff_GAT_3d(
            steps=self.steps, input_channels=self.input_channels, freeze_backbone=self.free_backbone,
            n_layers=self.n_layers, backbone=self.backbone, t_channels=3, architecture=self.architecture,
            gemm_mode=self.gemm_mode, attn_mode=self.attn_mode,
        )

    def forward(self, xy_pos, time, patch_rgb, edge_index, batch) -> Any:
        return self.model(xy_pos, time, patch_rgb, edge_index, batch)

    def forward_with_feats(self, xy_pos: Tensor, time: Tensor, edge_index: Tensor, pcd_feats: Tensor, batch,
                           return_attentions=False) -> Any:
        return self.model.forward_with_feats(xy_pos, time, edge_index, pcd_feats, batch)  # :369-385

    def pcd_features(self, pcd):
        return self.model.pcd_features(pcd)

    def _features_from_cond(self, cond):
        if cond.dim() == 2 and cond.shape[1] == self.model.combined_features_dim - 64:
            return cond
        return self.pcd_features(cond)

    @torch.no_grad()
    def p_sample_ddim(self, x, t, t_index, edge_index, pcd_feats, batch):  # :595-663
        # per-node t: coefficients gathered on the device, no host look at t (see GNN_Diffusion.p_sample_ddim)
        eng = self.model.engine_for(edge_index, pcd_feats, batch)
        return eng.ddim_step_t(x, t, self._pred_code(), 0.0, self._device_schedule()), None

    @torch.no_grad()
    def _p_sample(self, x, t, t_index, edge_index, sampling_func, pcd_feats, batch):
        return sampling_func(x, t, t_index, edge_index, pcd_feats, batch)

    @torch.no_grad()
    def p_sample_loop(self, shape, cond, edge_index, batch, generator: Optional[torch.Generator] = None):  # :688-731
        device = edge_index.device
        b = shape[0]
        img = torch.randn((b, 3), device=device, generator=generator) * self.noise_weight
        quat = torch.zeros((b, 4), device=device)
        quat[:, 0] = 1.0  # matrix_to_quaternion(eye(3)) == (1, 0, 0, 0)   (:704-709)
        img = torch.concat([quat, img], axis=1)
        imgs, attentions = [], []
        pcd_feats = self._features_from_cond(cond)
        eng = self.model.engine_for(edge_index, pcd_feats, batch)
        pred = self._pred_code()
        for i in list(reversed(range(0, self.steps, self.inference_ratio))):
            img = eng.ddim_step(img, self._step_coef(i, pred))
            attentions.append(None)
            imgs.append(img)
        return imgs, attentions
