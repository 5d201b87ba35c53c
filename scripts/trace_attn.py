"""Development aid: clock64 trace of one CTA of the dense attention kernel (critical-path analysis)."""
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
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
NCTA = 4096
buf = torch.zeros(128 + 8 * NCTA, dtype=torch.int64, device=dev)
buf[127] = NCTA
coef = mod._step_coef(290, mod._pred_code())
for _ in range(3): eng.ddim_step(x, coef)
lib.da_debug_trace(eng._h, C.c_void_p(buf.data_ptr()), layer)
eng.ddim_step(x, coef); torch.cuda.synchronize()
t = buf.cpu().tolist()   # trace of the LAST attention launch (last layer, C=144)
base = min(v for v in t[:127] if v > 0)
print("block | MMA: S(j+1) issue, p_full seen, v_full seen | softmax: wait start, S ready, P published   (clk rel.)")
for j in range(15):
    mm = [t[j*4+k]-base if t[j*4+k] else -1 for k in range(3)]
    sm = [t[64+j*4+k]-base if t[64+j*4+k] else -1 for k in range(3)]
    print(j, mm, sm)
e = [t[120 + k] - base if t[120 + k] else -1 for k in range(6)]
print("epilogue: last P published, last PV retired, residual scores done, staged rows landed, outputs stored, (chunk loop done):", e)

import collections
rec = buf[128:].view(NCTA, 8).cpu()
rec = rec[rec[:, 0] > 0]
t0 = int(rec[:, 0].min())
start, end, smid, cyc = (rec[:, 0] - t0).tolist(), (rec[:, 1] - t0).tolist(), rec[:, 2].tolist(), (rec[:, 7] - rec[:, 3]).tolist()
print(f"{len(start)} CTAs on {len(set(smid))} SMs; kernel span {max(end) / 1e3:.1f} us")
dur = [e - s for s, e in zip(start, end)]
import statistics
print("CTA duration us: min %.1f median %.1f mean %.1f max %.1f" % (min(dur) / 1e3, statistics.median(dur) / 1e3, statistics.mean(dur) / 1e3, max(dur) / 1e3))
print("CTA duration cycles: median %d   (=> SM clock %.2f GHz)" % (statistics.median(cyc), statistics.median(cyc) / statistics.median(dur)))
per_sm = collections.defaultdict(list)
for s, e, m in zip(start, end, smid): per_sm[m].append((s, e))
busy = []
for m, iv in per_sm.items():
    busy.append(sum(e - s for s, e in iv) / max(e for _, e in iv))
print("mean concurrent CTAs per SM: %.2f (min %.2f)   CTAs per SM: min %d max %d" % (statistics.mean(busy), min(busy), min(len(v) for v in per_sm.values()), max(len(v) for v in per_sm.values())))
for k in (0, 1, 300, 1000, 2000):
    if k < len(start): print("CTA", k, "sm", smid[k], "start %.1f us end %.1f us" % (start[k] / 1e3, end[k] / 1e3))

ph = torch.stack([rec[:, 4] - rec[:, 3], rec[:, 5] - rec[:, 4], rec[:, 6] - rec[:, 5], rec[:, 7] - rec[:, 6]], 1).float()
print("phases (cycles, median over CTAs): setup %d | Q parked %d | softmax loop %d | epilogue + teardown %d" % tuple(ph.median(0).values.tolist()))
print("phases (cycles, mean over CTAs):   setup %d | Q parked %d | softmax loop %d | epilogue + teardown %d" % tuple(ph.mean(0).tolist()))
late = ph[len(ph) // 2:]
print("second half of the grid (mean):    setup %d | Q parked %d | softmax loop %d | epilogue + teardown %d" % tuple(late.mean(0).tolist()))
