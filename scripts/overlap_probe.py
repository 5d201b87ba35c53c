"""Development aid: what slows the sampling loop down when the next batch is staged concurrently?"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
w = bench.WORKLOADS["c3_exphander60_v8"]
dev = torch.device("cuda", 0)
mod = bench.make_module(w, "bf16x3", "auto", dev)
M = w["n"] * w["B"]
ei_h, batch_h = bench.build_topology(w, dev, 0, mod)
g = torch.Generator().manual_seed(1)
feats_h = torch.randn(M, 1088, generator=g).pin_memory(); ei_h = ei_h.pin_memory(); batch_h = batch_h.pin_memory()
side = torch.cuda.Stream(dev)
nxt = mod.prefetch(feats_h, ei_h, batch_h)
ei_d, b_d, f_d = ei_h.to(dev), batch_h.to(dev), feats_h.to(dev)
spare_inputs = None

def run(mode):
    global nxt
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    imgs, _ = mod.p_sample_loop((M, 4), *nxt)
    e1.record()
    t1 = time.perf_counter()
    if mode == "prefetch":
        nxt = mod.prefetch(feats_h, ei_h, batch_h)
    elif mode == "h2d":
        with torch.cuda.stream(side):
            a = ei_h.to(dev, non_blocking=True); b = batch_h.to(dev, non_blocking=True); c = feats_h.to(dev, non_blocking=True)
        side.synchronize()
    elif mode == "plan":   # planning only, inputs already on the device
        sp = mod.model._spare
        with torch.cuda.stream(side):
            ext, nt, vi = mod.model.gnn_backbone.extend_graph(ei_d, b_d)
            sp.set_graph(ext, b_d, num_real=len(b_d), num_total=nt, virt_ids=vi)
            sp.set_features(f_d)
        side.synchronize()
    elif mode == "extend":
        with torch.cuda.stream(side):
            ext, nt, vi = mod.model.gnn_backbone.extend_graph(ei_d, b_d)
        side.synchronize()
    elif mode == "setgraph":
        sp = mod.model._spare
        with torch.cuda.stream(side):
            sp.set_graph(EXT[0], b_d, num_real=len(b_d), num_total=EXT[1], virt_ids=EXT[2])
        side.synchronize()
    elif mode == "setfeats":
        sp = mod.model._spare
        with torch.cuda.stream(side):
            sp.set_features(f_d)
        side.synchronize()
    elif mode == "memset":
        with torch.cuda.stream(side):
            for _ in range(4):
                Z.zero_()
        side.synchronize()
    elif mode == "sleep":
        time.sleep(0.015)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return f"{mode:9s} gpu loop {e0.elapsed_time(e1):7.2f} ms | host: enqueue {1e3*(t1-t0):6.2f} side {1e3*(t2-t1):6.2f} drain {1e3*(t3-t2):6.2f} total {1e3*(t3-t0):6.2f}"

EXT = mod.model.gnn_backbone.extend_graph(ei_d, b_d)
Z = torch.empty(150_000_000, dtype=torch.uint8, device=dev)
for mode in ["none", "none", "extend", "extend", "setgraph", "setgraph", "setfeats", "setfeats", "memset", "memset", "sleep", "none"]:
    if mode != "prefetch" and mod.model._prefetched is None:
        nxt = mod.prefetch(feats_h, ei_h, batch_h); torch.cuda.synchronize()
    print(run(mode))
    if mode != "prefetch":
        nxt = mod.prefetch(feats_h, ei_h, batch_h); torch.cuda.synchronize()
