"""Worker of tests/test_gpu_multirank.py (launched with torchrun, one rank per GPU, NCCL).

Checks the multi-GPU leg of the sampling path on real devices:
  * V = 0: the graph-sharded run (sharding.shard_batch -> per-rank fused DDIM steps -> ONE all-gather) equals the
    unsharded run on the whole batch BIT FOR BIT (graphs are independent: same kernels, same tiles, same order);
  * V = 8 (exophormer): the reference's virtual wiring couples the graphs of a batch, so each rank's result is held to
    the CPU oracle evaluated on THAT RANK'S sub-batch (1e-4);
  * training: gradients averaged with the NCCL all-reduce equal the single-process gradients of the whole batch.
Prints one line "MULTIRANK_OK ..." on rank 0 when everything holds.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import diffassemble_b200 as dab
    import oracle
    from common import reseed_parameters, rel_err, synth_graph_batch
    from diffassemble_b200 import sharding

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    msgs = []
    sizes = [100, 120, 100, 80]
    ei, batch = synth_graph_batch(sizes, kind="expander", degree="60%", seed=5)
    M = sum(sizes)
    g = torch.Generator().manual_seed(1)
    feats, x = torch.randn(M, 1088, generator=g), torch.randn(M, 4, generator=g)
    counts = []
    for r in range(world):
        g0, g1 = sharding.shard_bounds(len(sizes), world, r)
        counts.append(sum(sizes[g0:g1]))
    for V in (0, 8):
        ref = oracle.GNNDiffusionRef(steps=300, sampling="DDIM", rotation=True, architecture="exophormer", virt_nodes=V,
                                     model_mean_type=oracle.ModelMeanType.START_X, inference_ratio=10).eval()
        reseed_parameters(ref, 21 + V)
        mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", rotation=True, architecture="exophormer", virt_nodes=V,
                                model_mean_type=dab.ModelMeanType.START_X, inference_ratio=10, gemm_mode="bf16x3", attn_mode="auto")
        mod.load_state_dict(ref.state_dict(), strict=True)
        mod = mod.to(dev)
        ei_r, batch_r, (feats_r, x_r), (n0, n1) = sharding.shard_batch(ei, batch, [feats, x], world, rank)
        xs = x_r.to(dev)
        xo = x_r.clone()
        ei_d, b_d, f_d = ei_r.to(dev), batch_r.to(dev), feats_r.to(dev)
        for i in (290, 280, 0):
            t = torch.full((n1 - n0,), i, dtype=torch.long)
            xs, _ = mod.p_sample(xs, t.to(dev), i, cond=None, edge_index=ei_d, patch_feats=f_d, batch=b_d)
            if V > 0:
                with torch.no_grad():
                    xo, _ = ref.p_sample(xo, t, i, edge_index=ei_r, patch_feats=feats_r, batch=batch_r)
                e = rel_err(xs, xo)
                if not e < 1e-4:
                    ok = False
                msgs.append(f"V={V} rank {rank} t={i}: vs oracle on the sub-batch {e:.2e}")
                xs = xo.to(dev)   # teacher forcing
        full = sharding.gather_poses(xs, counts)          # the ONE collective
        if V == 0:
            mod.model.invalidate()
            xu = x.to(dev)
            ei_u, b_u, f_u = ei.to(dev), batch.to(dev), feats.to(dev)
            for i in (290, 280, 0):
                t = torch.full((M,), i, dtype=torch.long, device=dev)
                xu, _ = mod.p_sample(xu, t, i, cond=None, edge_index=ei_u, patch_feats=f_u, batch=b_u)
            same = torch.equal(full, xu)
            ok = ok and same
            msgs.append(f"V=0 rank {rank}: sharded + gathered == unsharded bit for bit: {same} (max diff {(full - xu).abs().max().item():.2e})")
    # ---- training leg: NCCL gradient all-reduce == whole-batch gradients -------------------------------------
    from diffassemble_b200.training import allreduce_gradients

    sizes_t = [64, 64, 64, 64]
    ei_t, batch_t = synth_graph_batch(sizes_t)
    Mt = sum(sizes_t)
    g = torch.Generator().manual_seed(2)
    feats_t, x0_t, noise_t = torch.randn(Mt, 1088, generator=g), torch.rand(Mt, 4, generator=g) * 2 - 1, torch.randn(Mt, 4, generator=g)
    tt = torch.tensor([5, 17, 250, 111])[batch_t]
    torch.manual_seed(0)
    mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", rotation=True, inference_ratio=10, model_mean_type=dab.ModelMeanType.START_X,
                            gemm_mode="bf16x3", attn_mode="auto")
    reseed_parameters(mod, 33)
    mod = mod.to(dev)
    params = [p for p in mod.parameters() if p.requires_grad]
    ei_r, batch_r, (f_r, x_r, nz_r, t_r), _ = sharding.shard_batch(ei_t, batch_t, [feats_t, x0_t, noise_t, tt], world, rank)
    loss = mod.p_losses(x_r.to(dev), t_r.to(dev), noise=nz_r.to(dev), loss_type="huber", cond=f_r.to(dev), edge_index=ei_r.to(dev),
                        batch=batch_r.to(dev))
    loss.backward()
    allreduce_gradients(params, world)
    sharded = [p.grad.clone() if p.grad is not None else None for p in params]
    for p in params:
        p.grad = None
    loss_u = mod.p_losses(x0_t.to(dev), tt.to(dev), noise=noise_t.to(dev), loss_type="huber", cond=feats_t.to(dev), edge_index=ei_t.to(dev),
                          batch=batch_t.to(dev))
    loss_u.backward()
    worst = 0.0
    for p, gs in zip(params, sharded):
        if p.grad is None or gs is None or p.grad.abs().max() < 1e-9:
            continue
        worst = max(worst, rel_err(gs, p.grad))
    # equal shard sizes: the mean of the per-shard mean losses is the whole-batch mean loss
    ok = ok and worst < 1e-3
    msgs.append(f"rank {rank}: all-reduced gradients vs whole-batch gradients, worst tensor {worst:.2e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    gathered = [None] * world
    dist.all_gather_object(gathered, msgs)
    if rank == 0:
        for ms in gathered:
            for m in ms:
                print(m)
        print("MULTIRANK_OK" if int(flag.item()) == 1 else "MULTIRANK_FAILED", f"world={world}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
