"""c2 (BASELINE configs[1]): 12x12 puzzle, 300-step DDPM sampling on one B200: eager loop vs whole-loop CUDA graph."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import diffassemble_b200 as dab
from diffassemble_b200 import topology
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mod = dab.GNN_Diffusion(steps=300, sampling="DDPM", rotation=True, noise_weight=1.0).to(dev)
n = 144
ei = topology.dense_edge_index(n).to(dev); batch = torch.zeros(n, dtype=torch.long, device=dev)
feats = torch.randn(n, 1088, device=dev)
for name, fn in [("eager", mod.p_sample_loop), ("graphed", mod.p_sample_loop_graphed)]:
    for _ in range(2): fn((n, 4), feats, ei, batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 5
    for _ in range(reps): imgs, _ = fn((n, 4), feats, ei, batch)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print(f"{name}: 300-step DDPM loop {dt*1e3:.2f} ms  -> {300/dt:.0f} denoising-steps/s ({dt/300*1e6:.1f} us/step)")
