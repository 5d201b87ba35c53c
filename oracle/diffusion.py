"""Oracle restatement of the diffusion schedules and samplers (test infrastructure).

* ``GNNDiffusionRef``   follows ``puzzle_diff/model/spatial_diffusion.py:130-176,
  219-357, 371-430, 485-699`` (2D puzzles; DDPM + DDIM).
* ``GNNDiffusion3dRef`` follows ``puzzle_diff/model/spatial_diffusion_3d_test_double_diffusion.py:
  177-186, 228-345, 575-737`` (R^3 + SO(3) DDIM).

Differences from the reference, all deliberate and documented in SURVEY.md section 0:
``p_sample_ddpm`` returns ``(x_prev, attentions)`` like ``p_sample_ddim`` does (the
shipped code returns a bare tensor which ``p_sample_loop`` then fails to unpack);
the encoders are out of scope, so ``cond`` is ignored and node features are passed
to ``p_sample_loop`` directly; every sampler accepts an optional ``noise=`` so tests
can teacher-force identical draws (when omitted the reference's own
``torch.randn_like`` call is made).
"""
import enum

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .eff_gat import EffGAT3dRef, EffGATRef
from .so3 import matrix_to_quaternion, quaternion_to_matrix, so3_scale


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelScheduler(enum.Enum):
    LINEAR = enum.auto()
    COSINE = enum.auto()
    COSINE_DISCRETE = enum.auto()


def cosine_discrete_beta_schedule(timesteps, s=0.08):  # spatial_diffusion.py:129-139
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps)
    alphas_cumprod = lambda t: torch.cos(((t / timesteps) + s) / (1 + s) + np.pi / 2)  # noqa: E731
    betas = 1 - alphas_cumprod(t + 1) / alphas_cumprod(t)
    return torch.clip(betas, 0.0001, 0.9999)


def cosine_beta_schedule(timesteps, s=0.08):  # :142-151
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps)
    alphas_cumprod = torch.cos(((x / timesteps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.clip(betas, 0.0001, 0.9999)


def linear_beta_schedule(timesteps):  # :154-157
    return torch.linspace(0.0001, 0.02, timesteps)


def extract(a, t, x_shape=None):  # :173-176
    out = a.gather(-1, t)
    return out[:, None]


def extract_rot(a, t, x_shape):  # ..._double_diffusion.py:183-186
    b = t.shape[0]
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


class _ScheduleMixin:
    def _register_schedule(self, steps, scheduler):
        # spatial_diffusion.py:282-321
        betas = {
            ModelScheduler.LINEAR: linear_beta_schedule,
            ModelScheduler.COSINE: cosine_beta_schedule,
            ModelScheduler.COSINE_DISCRETE: cosine_discrete_beta_schedule,
        }[scheduler](timesteps=steps)
        self.register_buffer("betas", betas)
        self.register_buffer("alphas", 1.0 - self.betas)
        self.register_buffer("alphas_cumprod", torch.cumprod(self.alphas, axis=0))
        self.register_buffer("alphas_cumprod_prev", F.pad(self.alphas_cumprod[:-1], (1, 0), value=1.0))
        self.register_buffer("sqrt_recip_alphas", torch.sqrt(1.0 / self.alphas))
        self.register_buffer("sqrt_alphas_cumprod", torch.sqrt(self.alphas_cumprod))
        self.register_buffer("sqrt_recip_alphas_cumprod", torch.sqrt(1.0 / self.alphas_cumprod))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", torch.sqrt(1.0 / self.alphas_cumprod - 1))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - self.alphas_cumprod))
        self.register_buffer(
            "posterior_variance", self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        )

    def _get_variance(self, timestep, prev_timestep):  # :528-546
        alpha_prod_t = extract(self.alphas_cumprod, timestep)
        alpha_prod_t_prev = (
            extract(self.alphas_cumprod, prev_timestep) if (prev_timestep >= 0).all() else alpha_prod_t * 0 + 1
        )
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        return (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):  # :629-632
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - pred_xstart) / extract(
            self.sqrt_recipm1_alphas_cumprod, t, x_t.shape
        )


class GNNDiffusionRef(nn.Module, _ScheduleMixin):
    def __init__(
        self,
        steps=600,
        inference_ratio=1,
        sampling="DDPM",
        classifier_free_prob=0,
        classifier_free_w=0,
        noise_weight=0.0,
        rotation=False,
        model_mean_type=ModelMeanType.EPSILON,
        input_channels=2,
        output_channels=2,
        scheduler=ModelScheduler.LINEAR,
        backbone="efficientnet_b0",
        n_layers=4,
        architecture="transformer",
        virt_nodes=4,
    ):
        super().__init__()
        self.model_mean_type = model_mean_type
        self.classifier_free_prob = classifier_free_prob
        self.classifier_free_w = classifier_free_w
        self.noise_weight = noise_weight
        self.rotation = rotation
        self.virt_nodes = virt_nodes
        self.inference_ratio = inference_ratio
        self.sampling = sampling
        self.eta = {"DDPM": 1, "DDIM": 0}[sampling]  # :264-278
        self._register_schedule(steps, scheduler)
        self.steps = steps
        extra = 2 if rotation else 0  # :334-345
        self.model = EffGATRef(
            steps=steps,
            input_channels=input_channels + extra,
            output_channels=output_channels + extra,
            model=backbone,
            architecture=architecture,
            n_layers=n_layers,
            virt_nodes=virt_nodes,
        )

    def forward_with_feats(self, xy_pos, time, patch_rgb, edge_index, patch_feats, batch, return_attentions=False):
        out, attentions = self.model.forward_with_feats(xy_pos, time, patch_rgb, edge_index, patch_feats, batch)
        if return_attentions:
            return out, attentions
        return out

    def q_sample(self, x_start, t, noise=None):  # :421-430
        if noise is None:
            noise = torch.randn_like(x_start)
        return (
            extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
            + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise
        )

    def p_losses(self, x_start, t, noise=None, loss_type="l1", edge_index=None, patch_feats=None, batch=None):
        # :432-483
        if noise is None:
            noise = torch.randn_like(x_start)
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        if self.steps == 1:
            x_noisy = torch.zeros_like(x_noisy)
        prediction = self.forward_with_feats(x_noisy, t, None, edge_index, patch_feats=patch_feats, batch=batch)
        target = {ModelMeanType.START_X: x_start, ModelMeanType.EPSILON: noise}[self.model_mean_type]
        if loss_type == "l1":
            return F.l1_loss(target, prediction)
        if loss_type == "l2":
            return F.mse_loss(target, prediction)
        if loss_type == "huber":
            return F.smooth_l1_loss(target, prediction)
        raise NotImplementedError()

    @torch.no_grad()
    def p_sample_ddpm(self, x, t, t_index, cond, edge_index, patch_feats, batch, noise=None):
        # :485-510
        betas_t = extract(self.betas, t, x.shape)
        sqrt_one_minus_alphas_cumprod_t = extract(self.sqrt_one_minus_alphas_cumprod, t, x.shape)
        sqrt_recip_alphas_t = extract(self.sqrt_recip_alphas, t, x.shape)
        model_out, attentions = self.forward_with_feats(
            x, t, cond, edge_index, patch_feats=patch_feats, batch=batch, return_attentions=True
        )
        model_mean = sqrt_recip_alphas_t * (x - betas_t * model_out / sqrt_one_minus_alphas_cumprod_t)
        if t_index == 0:
            return model_mean, attentions
        posterior_variance_t = extract(self.posterior_variance, t, x.shape)
        if noise is None:
            noise = torch.randn_like(x)
        return model_mean + torch.sqrt(posterior_variance_t) * noise, attentions

    @torch.no_grad()
    def p_sample_ddim(self, x, t, t_index, cond, edge_index, patch_feats, batch, noise=None):
        # :548-627
        prev_timestep = t - self.inference_ratio
        eta = self.eta
        alpha_prod = extract(self.alphas_cumprod, t, x.shape)
        if (prev_timestep >= 0).all():
            alpha_prod_prev = extract(self.alphas_cumprod, prev_timestep, x.shape)
        else:
            alpha_prod_prev = alpha_prod * 0 + 1
        beta = 1 - alpha_prod
        if self.classifier_free_prob > 0.0:
            model_output_cond, attentions = self.forward_with_feats(
                x, t, cond, edge_index, patch_feats=patch_feats, batch=batch, return_attentions=True
            )
            model_output_uncond = self.forward_with_feats(
                x, t, cond, edge_index, patch_feats=torch.zeros_like(patch_feats), batch=batch
            )
            model_output = (1 + self.classifier_free_w) * model_output_cond - self.classifier_free_w * model_output_uncond
        else:
            model_output, attentions = self.forward_with_feats(
                x, t, cond, edge_index, patch_feats=patch_feats, batch=batch, return_attentions=True
            )
        x_0 = {
            ModelMeanType.EPSILON: (x - beta**0.5 * model_output) / alpha_prod**0.5,
            ModelMeanType.START_X: model_output,
        }[self.model_mean_type]
        eps = self._predict_eps_from_xstart(x, t, x_0)
        variance = self._get_variance(t, prev_timestep)
        std_eta = eta * variance**0.5
        pred_sample_direction = (1 - alpha_prod_prev - std_eta**2) ** (0.5) * eps
        prev_sample = alpha_prod_prev ** (0.5) * x_0 + pred_sample_direction
        if eta > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, dtype=model_output.dtype)
            prev_sample = prev_sample + std_eta * noise
        return prev_sample, attentions

    def p_sample(self, x, t, t_index, cond=None, edge_index=None, patch_feats=None, batch=None, noise=None):
        fn = self.p_sample_ddpm if self.sampling == "DDPM" else self.p_sample_ddim
        return fn(x, t, t_index, cond, edge_index, patch_feats, batch, noise=noise)

    @torch.no_grad()
    def p_sample_loop(self, shape, patch_feats, edge_index, batch, generator=None, keep_attentions=False):
        # :635-676; ``patch_feats`` replaces ``visual_features(cond)`` (encoder out of scope)
        b = shape[0]
        img = torch.randn(shape, generator=generator) * self.noise_weight
        imgs, attentions = [], []
        for i in list(reversed(range(0, self.steps, self.inference_ratio))):
            needs_noise = (self.sampling == "DDPM" and i != 0) or (self.sampling == "DDIM" and self.eta > 0)
            noise = torch.randn(shape, generator=generator) if needs_noise else None
            img, atts = self.p_sample(
                img,
                torch.full((b,), i, dtype=torch.long),
                i,
                cond=None,
                edge_index=edge_index,
                patch_feats=patch_feats,
                batch=batch,
                noise=noise,
            )
            attentions.append(atts if keep_attentions else None)
            imgs.append(img)
        return imgs, attentions


class GNNDiffusion3dRef(nn.Module, _ScheduleMixin):
    def __init__(
        self,
        steps=600,
        inference_ratio=1,
        sampling="DDIM",
        noise_weight=0.0,
        model_mean_type=ModelMeanType.EPSILON,
        input_channels=7,
        scheduler=ModelScheduler.LINEAR,
        n_layers=4,
        backbone="pointnet",
        architecture="transformer",
    ):
        super().__init__()
        assert sampling == "DDIM"  # the only sampler bound, ..._double_diffusion.py:275-281
        self.model_mean_type = model_mean_type
        self.noise_weight = noise_weight
        self.inference_ratio = inference_ratio
        self.eta = 0
        self._register_schedule(steps, scheduler)
        self.steps = steps
        self.model = EffGAT3dRef(
            steps=steps, input_channels=input_channels, n_layers=n_layers, backbone=backbone, architecture=architecture
        )

    def forward_with_feats(self, xy_pos, time, edge_index, pcd_feats, batch, return_attentions=False):
        return self.model.forward_with_feats(xy_pos, time, edge_index, pcd_feats, batch)

    def _predict_eps_from_xstart_rot(self, x_t, t, pred_xstart):  # :670-685
        x_t_term = so3_scale(
            quaternion_to_matrix(x_t),
            (
                extract_rot(self.sqrt_recip_alphas_cumprod, t, t.shape)
                / extract_rot(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape).flatten()
            ),
        )
        pred_xstart = so3_scale(
            quaternion_to_matrix(pred_xstart), 1 / extract_rot(self.sqrt_recipm1_alphas_cumprod, t, t.shape)
        )
        return x_t_term @ pred_xstart.transpose(-1, -2)

    @torch.no_grad()
    def p_sample_ddim(self, x, t, t_index, edge_index, pcd_feats, batch):  # :595-663
        prev_timestep = t - self.inference_ratio
        alpha_prod = extract(self.alphas_cumprod, t, x.shape)
        if (prev_timestep >= 0).all():
            alpha_prod_prev = extract(self.alphas_cumprod, prev_timestep, x.shape)
        else:
            alpha_prod_prev = alpha_prod * 0 + 1
        beta = 1 - alpha_prod
        model_output, attentions = self.forward_with_feats(x, t, edge_index, pcd_feats=pcd_feats, batch=batch)
        x_0 = {
            ModelMeanType.EPSILON: (x - beta**0.5 * model_output) / alpha_prod**0.5,
            ModelMeanType.START_X: model_output,
        }[self.model_mean_type]
        x_0_tr, x_0_r = x_0[:, 4:], x_0[:, :4]
        x_tr, x_quater = x[:, 4:], x[:, :4]
        eps_tr = self._predict_eps_from_xstart(x_tr, t, x_0_tr)
        eps_rot = matrix_to_quaternion(self._predict_eps_from_xstart_rot(x_quater, t, x_0_r))
        pred_sample_direction_tr = (1 - alpha_prod_prev) ** (0.5) * eps_tr
        pred_sample_direction_rot = so3_scale(quaternion_to_matrix(eps_rot), ((1 - alpha_prod_prev) ** (0.5)).view(-1))
        prev_sample_tr = alpha_prod_prev ** (0.5) * x_0_tr + pred_sample_direction_tr
        prev_sample_r = matrix_to_quaternion(
            so3_scale(quaternion_to_matrix(x_0_r), (alpha_prod_prev ** (0.5)).view(-1)) @ pred_sample_direction_rot
        )
        return torch.concat([prev_sample_r, prev_sample_tr], axis=1), attentions

    def p_sample(self, x, t, t_index, edge_index=None, pcd_feats=None, batch=None):
        return self.p_sample_ddim(x, t, t_index, edge_index, pcd_feats, batch)

    @torch.no_grad()
    def p_sample_loop(self, shape, pcd_feats, edge_index, batch, generator=None):  # :688-731
        b = shape[0]
        img = torch.randn((b, 3), generator=generator) * self.noise_weight
        x = matrix_to_quaternion(torch.eye(3).repeat(b, 1, 1))
        img = torch.concat([x, img], axis=1)
        imgs, attentions = [], []
        for i in list(reversed(range(0, self.steps, self.inference_ratio))):
            img, atts = self.p_sample(
                img, torch.full((b,), i, dtype=torch.long), i, edge_index=edge_index, pcd_feats=pcd_feats, batch=batch
            )
            attentions.append(None)
            imgs.append(img)
        return imgs, attentions
