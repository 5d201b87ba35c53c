"""Development aid: kernel-level breakdown of one c5 training step (torch profiler, CUDA activities)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import diffassemble_b200 as dab
from diffassemble_b200 import topology
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
B, n = 64, 144
torch.manual_seed(0)
mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", rotation=True, inference_ratio=10, model_mean_type=dab.ModelMeanType.START_X).to(dev)
ei, batch = topology.batch_graphs([topology.dense_edge_index(n)] * B, [n] * B)
ei, batch = ei.to(dev), batch.to(dev)
M = B * n
feats, x0 = torch.randn(M, 1088, device=dev), torch.rand(M, 4, device=dev) * 2 - 1
opt = mod.configure_optimizers()
def step():
    t = torch.randint(0, 300, (B,), device=dev)[batch]
    opt.zero_grad(set_to_none=True)
    loss = mod.p_losses(x0, t, loss_type="huber", cond=feats, edge_index=ei, batch=batch)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
