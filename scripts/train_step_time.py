"""c5 (BASELINE configs[4]): one training step (forward + backward + Adafactor) on 12x12 puzzles, per GPU."""
import sys, time, os, torch
sys.path.insert(0, '/root/repo')
import diffassemble_b200 as dab
from diffassemble_b200 import topology
import oracle
dev = torch.device("cuda", 0)
B, n = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 144
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
    return loss
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter(); K = 10
for _ in range(K): loss = step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
print(f"GPU: {B} x {n}-node graphs: {dt*1e3:.1f} ms / training step ({B/dt:.0f} graphs/s), loss {loss.item():.4f}")
# CPU oracle (autograd + the same Adafactor) on a bounded sample
Bc = 4
torch.set_num_threads(os.cpu_count())
ref = oracle.GNNDiffusionRef(steps=300, sampling="DDIM", rotation=True, inference_ratio=10, model_mean_type=oracle.ModelMeanType.START_X)
from transformers.optimization import Adafactor
opt_c = Adafactor(ref.parameters())
eic, batchc = oracle.batch_graphs([oracle.dense_edge_index(n)] * Bc, [n] * Bc)
fc, xc = torch.randn(Bc * n, 1088), torch.rand(Bc * n, 4)
def cstep():
    t = torch.randint(0, 300, (Bc,))[batchc]
    opt_c.zero_grad(); l = ref.p_losses(xc, t, loss_type="huber", edge_index=eic, patch_feats=fc, batch=batchc); l.backward(); opt_c.step()
cstep(); t0 = time.perf_counter(); cstep(); cstep(); dtc = (time.perf_counter() - t0) / 2
print(f"CPU oracle ({os.cpu_count()} threads): {Bc} graphs: {dtc*1e3:.0f} ms / step ({Bc/dtc:.1f} graphs/s)")
# phase breakdown of the GPU step
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
t = torch.randint(0, 300, (B,), device=dev)[batch]
opt.zero_grad(set_to_none=True)
loss, t_f = timed(lambda: mod.p_losses(x0, t, loss_type="huber", cond=feats, edge_index=ei, batch=batch))
_, t_b = timed(lambda: loss.backward())
_, t_o = timed(lambda: opt.step())
print(f"breakdown: forward {t_f:.1f} ms, backward {t_b:.1f} ms, Adafactor {t_o:.1f} ms")
