import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
w = bench.WORKLOADS["c3_exphander60_v8"]
dev = torch.device("cuda", 0)
mod = bench.make_module(w, "bf16x3", "auto", dev)
M = w["n"] * w["B"]
t0 = time.perf_counter(); ei_h, batch_h = bench.build_topology(w, dev, 0, mod); print("topology host", time.perf_counter() - t0)
g = torch.Generator().manual_seed(1)
feats_h = torch.randn(M, 1088, generator=g).pin_memory(); ei_h = ei_h.pin_memory(); batch_h = batch_h.pin_memory()
def sync(): torch.cuda.synchronize()
for it in range(4):
    mod.model.invalidate()
    sync(); t0 = time.perf_counter()
    ei = ei_h.to(dev, non_blocking=True); batch = batch_h.to(dev, non_blocking=True); feats = feats_h.to(dev, non_blocking=True)
    sync(); t1 = time.perf_counter()
    eng = mod.model._get_engine(dev)
    sync(); t2 = time.perf_counter()
    ext, num_total, virt_ids = mod.model.gnn_backbone.extend_graph(ei, batch)
    sync(); t3 = time.perf_counter()
    eng.set_graph(ext, batch, num_real=len(batch), num_total=num_total, virt_ids=virt_ids)
    sync(); t4 = time.perf_counter()
    eng.set_features(feats)
    sync(); t5 = time.perf_counter()
    mod.model._graph_key = (mod.model._tensor_key(ei), mod.model._tensor_key(batch)); mod.model._feats_key = mod.model._tensor_key(feats)
    imgs, _ = mod.p_sample_loop((M, 4), feats, ei, batch)
    sync(); t6 = time.perf_counter()
    print(f"iter {it}: h2d {t1-t0:.4f} engine/weights {t2-t1:.4f} extend {t3-t2:.4f} set_graph {t4-t3:.4f} set_feats {t5-t4:.4f} loop {t6-t5:.4f} total {t6-t0:.4f}")
