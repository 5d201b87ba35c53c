"""Development aid: clock64 trace of CTA 0 of the persistent hidden-layer kernel (attn_hidden.cu, DBG variant)."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
import bench
w = bench.WORKLOADS["c3_exphander60_v8"]
dev = torch.device("cuda", 0)
mod = bench.make_module(w, "bf16x3", "auto", dev)
M = w["n"] * w["B"]
spec, _, _, _, _ = bench.host_batch(w, 1, 0, "strong")
ei, batch = spec.build(dev)
feats, x = torch.randn(M, 1088, device=dev), torch.randn(M, 4, device=dev)
eng = mod.model.engine_for(ei, feats, batch)
lib = eng._lib
lib.da_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 1
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
coef = mod._step_coef(290, mod._pred_code())
for _ in range(3): eng.ddim_step(x, coef)
lib.da_debug_trace(eng._h, C.c_void_p(buf.data_ptr()), layer)
eng.ddim_step(x, coef); torch.cuda.synchronize()
lib.da_debug_trace(eng._h, None, -1)
t = buf.cpu().tolist()
base = min(v for v in t if v > 0)
for s in range(2):
    for it in range(8):
        q = [t[2048 + (s * 8 + it) * 4 + k] - base for k in range(4)]
        if q[0] < 0: continue
        print(f"stream {s} item {it}: loop end {q[0]}  Q(next) parked +{q[1]-q[0]}  O complete +{q[2]-q[1]}  epilogue done +{q[3]-q[2]}")
        if it not in (0, 3): continue
        print("  blk | softmax: wait start, +S ready, +S in regs, +P published | MMA: S(j+1) issued, p_full seen, PV issued (rel. to softmax wait start)")
        for j in range(16):
            r = t[((s * 8 + it) * 16 + j) * 8:((s * 8 + it) * 16 + j) * 8 + 7]
            if r[0] == 0: continue
            print(f"  {j:3d} | {r[0]-base:7d} +{r[1]-r[0]:5d} +{r[2]-r[1]:5d} +{r[3]-r[2]:5d} | {r[4]-r[0]:6d} {r[5]-r[0]:6d} {r[6]-r[0]:6d}")
