"""Development aid: what the planner makes of one workload (block lists, reordering) and how long da_set_graph takes."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3_exphander60_v8"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = dict(bench.WORKLOADS[name]); w["B"] = B
dev = torch.device("cuda", 0)
mod = bench.make_module(w, "bf16x3", "auto", dev)
spec, feats_h, x_h, g0, Bl = bench.host_batch(w, 1, 0, "strong")
ei, batch = spec.build(dev)
feats = feats_h.to(dev)
for rep in range(3):
    mod.model.invalidate()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng = mod.model.engine_for(ei, feats, batch)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"bind graph + features: {dt * 1e3:.2f} ms", eng.plan_info(), eng.graph_stats())
