"""Generates the golden fixtures in this directory from the CPU oracle.

The reference ships no golden vectors for this path and cannot be imported in the build
image, so these fixtures pin the ORACLE (they detect drift of the
restatement and give the GPU tests a torch-version-independent target).  Weights are NOT
stored (12.8 MB per model): they are regenerated with ``torch.manual_seed(seed)`` and
checked against the stored checksum.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from common import synth_graph_batch  # noqa: E402

HERE = Path(__file__).resolve().parent


def weight_checksum(module):
    return float(sum(v.double().abs().sum() for k, v in sorted(module.state_dict().items()) if v.is_floating_point()))


def case_2d(name, sizes, architecture, virt_nodes, sampling, mean_type, ratio, kind="dense", degree="60%", loop_steps=None, T=300):
    torch.manual_seed(0)
    ref = oracle.GNNDiffusionRef(steps=T, sampling=sampling, rotation=True, architecture=architecture,
                                 virt_nodes=virt_nodes, model_mean_type=oracle.ModelMeanType[mean_type],
                                 inference_ratio=ratio)
    ref.eval()
    ei, batch = synth_graph_batch(sizes, kind=kind, degree=degree, seed=0)
    M = len(batch)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(M, 1088, generator=g)
    x = torch.randn(M, 4, generator=g)
    t = torch.full((M,), T - 1, dtype=torch.long)
    with torch.no_grad():
        out, atts = ref.model.forward_with_feats(x, t, None, ei, feats, batch)
        # one teacher-forced sampler step at a mid timestep, with noise
        ti = (T // 2 // ratio) * ratio
        noise = torch.randn(M, 4, generator=g)
        tt = torch.full((M,), ti, dtype=torch.long)
        step_out, _ = ref.p_sample(x, tt, ti, edge_index=ei, patch_feats=feats, batch=batch, noise=noise)
    d = dict(sizes=sizes, architecture=architecture, virt_nodes=virt_nodes, sampling=sampling, mean_type=mean_type,
             ratio=ratio, T=T, edge_index=ei, batch=batch, feats=feats, x=x, t=t, out=out, alpha_last=atts[-1][1],
             step_t=ti, step_noise=noise, step_out=step_out, weight_checksum=weight_checksum(ref))
    if loop_steps:
        # short full loop (T = loop_steps) to pin trajectory-level behaviour
        torch.manual_seed(0)
        ref2 = oracle.GNNDiffusionRef(steps=loop_steps, sampling=sampling, rotation=True, architecture=architecture,
                                      virt_nodes=virt_nodes, model_mean_type=oracle.ModelMeanType[mean_type],
                                      inference_ratio=1, noise_weight=1.0)
        ref2.eval()
        gen = torch.Generator().manual_seed(2)
        imgs, _ = ref2.p_sample_loop((M, 4), feats, ei, batch, generator=gen)
        d.update(loop_T=loop_steps, loop_final=imgs[-1], loop_first=imgs[0], loop_weight_checksum=weight_checksum(ref2))
    torch.save(d, HERE / f"{name}.pt")
    print(name, "out", tuple(out.shape), "E", ei.shape[1])


def case_3d(name, sizes, T=300, ratio=10):
    torch.manual_seed(0)
    ref = oracle.GNNDiffusion3dRef(steps=T, backbone="pointnet", inference_ratio=ratio,
                                   model_mean_type=oracle.ModelMeanType.START_X)
    ref.eval()
    ei, batch = synth_graph_batch(sizes, kind="dense")
    M = len(batch)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(M, 128, generator=g)
    q = torch.nn.functional.normalize(torch.randn(M, 4, generator=g), dim=-1)
    x = torch.cat([q, torch.randn(M, 3, generator=g)], 1)
    ti = 150
    t = torch.full((M,), ti, dtype=torch.long)
    with torch.no_grad():
        out, _ = ref.model.forward_with_feats(x, t, ei, feats, batch)
        step_out, _ = ref.p_sample(x, t, ti, edge_index=ei, pcd_feats=feats, batch=batch)
        gen = torch.Generator().manual_seed(2)
        ref.noise_weight = 1.0
        imgs, _ = ref.p_sample_loop((M, 7), feats, ei, batch, generator=gen)
    d = dict(sizes=sizes, T=T, ratio=ratio, edge_index=ei, batch=batch, feats=feats, x=x, t=t, out=out, step_t=ti,
             step_out=step_out, loop_final=imgs[-1], loop_first=imgs[0], weight_checksum=weight_checksum(ref))
    torch.save(d, HERE / f"{name}.pt")
    print(name, "out", tuple(out.shape))


if __name__ == "__main__":
    # c1: 6x6 dense, single forward (BASELINE configs[0]) + DDPM step + 8-step DDPM loop
    case_2d("c1_dense36_ddpm", [36], "transformer", 0, "DDPM", "EPSILON", 1, loop_steps=8)
    # ragged dense batch, DDIM x0-prediction (the shipped launch-script setting)
    case_2d("dense_ragged_ddim", [16, 25, 9], "transformer", 0, "DDIM", "START_X", 10, loop_steps=6)
    # exophormer: sparse expander + virtual nodes, two graphs (exercises the cross-graph wiring)
    case_2d("exph_2x64_ddim", [64, 64], "exophormer", 4, "DDIM", "START_X", 10, kind="expander", degree="60%", loop_steps=4)
    # c4-like: ragged 3D fragments, SE(3) head + SO(3) DDIM
    case_3d("se3_ragged", [2, 5, 20, 11, 7])
