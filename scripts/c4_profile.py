"""Development aid: per-launch-site times of one 3-D DDIM step (c4: 64 ragged fragment graphs) and of c2."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench, oracle
import diffassemble_b200 as dab
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c4":
    sizes, ei, batch, feats, x = bench.c4_inputs(64)
    mod = dab.GNN_Diffusion_3d(steps=300, sampling="DDIM", backbone="pointnet", inference_ratio=10, model_mean_type=dab.ModelMeanType.START_X).to(dev)
else:
    from diffassemble_b200 import topology
    n = 144
    ei, batch = topology.dense_edge_index(n), torch.zeros(n, dtype=torch.long)
    feats, x = torch.randn(n, 1088), torch.randn(n, 4)
    mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", rotation=True, inference_ratio=10, model_mean_type=dab.ModelMeanType.START_X).to(dev)
ei, batch, feats, x = ei.to(dev), batch.to(dev), feats.to(dev), x.to(dev)
eng = mod.model.engine_for(ei, feats, batch)
coef = mod._step_coef(290, mod._pred_code())
for _ in range(5): eng.ddim_step(x, coef)
eng.set_profiling(True); eng.get_profile(reset=True)
for _ in range(20): eng.ddim_step(x, coef)
torch.cuda.synchronize()
p = eng.get_profile(reset=True)
tot = 0
for k, v in p.items():
    if v["launches"]:
        print(f"{k:22s} {v['launches'] / 20:5.1f} launches/step  {v['ms'] / v['launches'] * 1e3:8.1f} us each  {v['ms'] / 20 * 1e3:8.1f} us/step")
        tot += v["ms"] / 20 * 1e3
print("sum", round(tot, 1), "us/step;", "nodes", x.shape[0], "edges", ei.shape[1], eng.graph_stats(), eng.plan_info())
