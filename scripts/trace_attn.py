"""Development aid: clock64 trace of one CTA of the dense attention kernel (critical-path analysis)."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
import bench
w = bench.WORKLOADS["c3_exphander60_v8"]
dev = torch.device("cuda", 0)
mod = bench.make_module(w, "bf16x3", "auto", dev)
M = w["n"] * w["B"]
ei, batch = bench.build_topology(w, dev, 0, mod)
ei, batch = ei.to(dev), batch.to(dev)
feats, x = torch.randn(M, 1088, device=dev), torch.randn(M, 4, device=dev)
eng = mod.model.engine_for(ei, feats, batch)
lib = eng._lib
lib.da_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
buf = torch.zeros(128, dtype=torch.int64, device=dev)
coef = mod._step_coef(290, mod._pred_code())
for _ in range(3): eng.ddim_step(x, coef)
lib.da_debug_trace(eng._h, C.c_void_p(buf.data_ptr()), layer)
eng.ddim_step(x, coef); torch.cuda.synchronize()
t = buf.cpu().tolist()   # trace of the LAST attention launch (last layer, C=144)
base = min(v for v in t if v > 0)
print("block | MMA: S(j+1) issue, p_full seen, v_full seen | softmax: wait start, S ready, P published   (clk rel.)")
for j in range(15):
    mm = [t[j*4+k]-base if t[j*4+k] else -1 for k in range(3)]
    sm = [t[64+j*4+k]-base if t[64+j*4+k] else -1 for k in range(3)]
    print(j, mm, sm)
e = [t[120 + k] - base if t[120 + k] else -1 for k in range(6)]
print("epilogue: last P published, last PV retired, residual scores done, staged rows landed, outputs stored, (chunk loop done):", e)
