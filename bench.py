#!/usr/bin/env python
"""Benchmark of the hot path: denoising steps/sec on 900-node (30x30) puzzle graphs.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A "step" is one pass of the hot path over one batch: one denoiser forward fused
with the sampler update for every graph of the batch.  `value` counts graph-steps (denoising steps
x graphs) per second over all ranks, inputs resident in HBM; `e2e` is the same metric through the
public `GNN_Diffusion.prefetch` / `p_sample_loop` API with pinned HOST buffers (uploads of features
and topology, and the per-step device->host read of x_t, inside the timed region).

Default workload (BASELINE.json configs[2]): 30x30 puzzles (900 nodes), Exphander 60 % sparse edges
(d = 539), architecture="exophormer" with 8 virtual nodes, a GLOBAL batch of 32 graphs sharded over
the ranks in contiguous blocks (`sharding.shard_batch`: 32 / 16 / 8 / 4 graphs per GPU at 1 / 2 / 4 /
8 GPUs, "scaling": "strong"; `--scaling weak` keeps 32 graphs per GPU), DDIM with x0-prediction,
T = 300, inference_ratio = 10 (the shipped launch-script setting), fp32 state, synthetic N(0,1)
patch features and seeded random-init weights.  The other workloads of SURVEY.md section 8(d)
(`--workload`): Exphander 20 % / 40 %, V = 0 / 4, the dense 900-node graph, 300-step DDPM, c2 (one
12x12 puzzle, 300-step DDPM, eager or whole-loop CUDA graph), c4 (64 ragged 3-D fragment graphs,
SE(3) head) and c5 (one training step on 64 12x12 puzzles).

`--impl reference` times the reference formulation on the host cores: the pure-torch edge-list
oracle (PyG / the reference package cannot be installed in this image), on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "denoising-steps/sec on 900-node (30x30) puzzle graphs"
UNIT = "graph-steps/s"


def _c3(degree, V, arch="exophormer", topo="expander", sampler="ddim"):
    return dict(kind="sample2d", n=900, B=32, topo=topo, degree=degree, arch=arch, V=V, sampler=sampler)


WORKLOADS = {
    # BASELINE.json configs[2] and the sweep of SURVEY.md section 8(d)
    "c3_exphander60_v8": _c3("60%", 8),
    "c3_exphander60_v4": _c3("60%", 4),
    "c3_exphander60_v0": _c3("60%", 0),
    "c3_exphander40_v8": _c3("40%", 8),
    "c3_exphander20_v8": _c3("20%", 8),
    "c3_dense": _c3(None, 0, arch="transformer", topo="dense"),
    "c3_exphander60_v8_ddpm300": _c3("60%", 8, sampler="ddpm"),
    # configs[1]: one 12x12 puzzle, 300-step DDPM (latency regime); *_graphed = whole loop as ONE CUDA graph
    "c2_dense144": dict(kind="sample2d", n=144, B=1, topo="dense", degree=None, arch="transformer", V=0, sampler="ddpm"),
    "c2_dense144_graphed": dict(kind="sample2d", n=144, B=1, topo="dense", degree=None, arch="transformer", V=0, sampler="ddpm",
                                graphed=True),
    # configs[3]: 64 ragged fragment graphs (2..20 nodes), PointNet-width features, SE(3) head, 3-D DDIM
    "c4_breakingbad64": dict(kind="sample3d", B=64),
    # configs[4]: one training step (forward + backward + Adafactor) on 64 12x12 puzzles
    "c5_train144": dict(kind="train", n=144, B=64),
}
T_STEPS, RATIO = 300, 10


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


# ---- work accounting (SURVEY.md section 8d / BASELINE.md section 2) -------------------------------------------
def flops_per_node(D=1152, C=4, Hm=128, hid=256):
    return 2 * (C * 16 + 16 * 32 + D * Hm + Hm * D + 4 * D * hid + 2 * (4 * hid * hid) + 4 * hid * D + D * 32 + 32 * C)


def flops_per_edge(D=1152, hid=256):
    return 4 * (3 * hid + D)


def bytes_per_node(D=1152, Dv=1088, C=4, hid=256):
    shc = 3 * hid + D
    return 4 * (Dv + C + C + 3 * D + 4 * shc + 2 * 3 * hid)


def step_work(M, E_tot, ms_per_step, pk, passes):
    sec = ms_per_step * 1e-3
    f_gemm, f_attn = M * flops_per_node(), E_tot * flops_per_edge()
    byt = bytes_per_node() * M + 16 * E_tot
    w = {"gflop_per_step": (f_gemm + f_attn) / 1e9, "compulsory_gb_per_step": byt / 1e9,
         "achieved_tflops_total": (f_gemm + f_attn) / sec / 1e12,
         "achieved_tflops_gemm": f_gemm / sec / 1e12, "achieved_tflops_attention": f_attn / sec / 1e12,
         "achieved_hbm_gbs_compulsory": byt / sec / 1e9}
    w["frac_tensor_peak"] = w["achieved_tflops_total"] / pk["tensor_sustained"]
    w["frac_tensor_peak_at_%d_passes" % passes] = w["achieved_tflops_total"] * passes / pk["tensor_sustained"]
    w["frac_hbm_peak"] = w["achieved_hbm_gbs_compulsory"] / pk["hbm"]
    w["step_roofline_frac"] = max(w["frac_tensor_peak"], w["frac_hbm_peak"])
    return w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin, self.t_end = None, None

    def begin(self):
        """Call right before the timed region: waits (<= 0.5 s) until nvidia-smi delivers rows, so that even a 40 ms region
        (the 4-graph share of the 8-GPU run) is sampled, then marks the start."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < 0.5:
            time.sleep(0.005)
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if self.t_begin is not None and self.t_end is not None:   # rows that arrived during the timed region (+ one period)
            rows = [r for r in self.rows if self.t_begin <= r[0] <= self.t_end + 0.02]
        for r in rows:
            r = r[1:]
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_dist(args):
    import torch.distributed as dist

    world, rank, local_rank = dist_env()
    if world != args.gpus and rank == 0 and world > 1:
        print(f"warning: WORLD_SIZE={world} != --gpus {args.gpus}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version / debug lines to stdout by default: keep stdout for the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    return world, rank, local_rank, device


def max_over_ranks(v, device, world):
    import torch.distributed as dist

    if world == 1:
        return v
    t = torch.tensor([v], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def graph_seeds(w, g0, g1):
    """Graph g of the GLOBAL batch always gets the same permutation, however the batch is sharded."""
    return [1000 + g for g in range(g0, g1)]


def make_module(w, gemm_mode, attn_mode, device):
    import diffassemble_b200 as dab

    torch.manual_seed(0)
    ddpm = w["sampler"] == "ddpm"
    mod = dab.GNN_Diffusion(steps=T_STEPS, sampling="DDPM" if ddpm else "DDIM", inference_ratio=1 if ddpm else RATIO, rotation=True,
                            noise_weight=1.0, model_mean_type=dab.ModelMeanType.EPSILON if ddpm else dab.ModelMeanType.START_X,
                            architecture=w["arch"], virt_nodes=w["V"], gemm_mode=gemm_mode, attn_mode=attn_mode)
    return mod.to(device)


def host_batch(w, world, rank, scaling):
    """This rank's share of the workload as HOST data: (topology spec, features, x_T, first graph, graphs on this rank).

    strong scaling: the global batch of w["B"] graphs is generated identically on every rank and cut with
    sharding.shard_bounds (the same contiguous blocks sharding.shard_batch produces from a collated batch);
    weak scaling: every rank owns w["B"] graphs of its own."""
    from diffassemble_b200 import sharding, topology

    n, B = w["n"], w["B"]
    if scaling == "strong":
        g0, g1 = sharding.shard_bounds(B, world, rank)
        gen = torch.Generator().manual_seed(1)
        feats_all = torch.randn(B * n, 1088, generator=gen)
        x_all = torch.randn(B * n, 4, generator=gen)
        feats, x = feats_all[g0 * n:g1 * n].clone(), x_all[g0 * n:g1 * n].clone()
    else:
        g0, g1 = rank * B, (rank + 1) * B
        gen = torch.Generator().manual_seed(1 + rank)
        feats, x = torch.randn(B * n, 1088, generator=gen), torch.randn(B * n, 4, generator=gen)
    Bl = g1 - g0
    if w["topo"] == "dense":
        spec = topology.DenseBatchSpec(n, Bl)
    else:
        spec = topology.ExpanderBatchSpec(torch.from_numpy(topology.expander_permutations(n, graph_seeds(w, g0, g1))).pin_memory(), w["degree"])
    return spec, feats.pin_memory(), x.pin_memory(), g0, Bl


def run_sample2d(args, w):
    import torch.distributed as dist

    world, rank, local_rank, device = init_dist(args)
    from diffassemble_b200 import sharding, topology

    mod = make_module(w, args.gemm, args.attn, device)
    ddpm = w["sampler"] == "ddpm"
    scaling = args.scaling if w["B"] >= 2 else "weak"
    spec, feats_h, x_h, g0, Bl = host_batch(w, world, rank, scaling)
    B_global = w["B"] if scaling == "strong" else w["B"] * world
    n = w["n"]
    M = Bl * n

    # ---- strong scaling goes through sharding.shard_batch once, on the collated global batch, and must agree with the
    # per-rank construction above (same graphs, same node ranges)
    if scaling == "strong" and world > 1 and w["topo"] != "dense":
        full = topology.ExpanderBatchSpec(topology.expander_permutations(n, graph_seeds(w, 0, w["B"])), w["degree"])
        ei_full, batch_full = full.build(device)
        ei_s, batch_s, _, (n0, n1) = sharding.shard_batch(ei_full, batch_full, [], world, rank)
        ei_l, batch_l = spec.build(device)
        assert (n0, n1) == (g0 * n, (g0 + Bl) * n) and torch.equal(ei_s, ei_l) and torch.equal(batch_s, batch_l), "shard_batch mismatch"
        del ei_full, batch_full, ei_s, batch_s, ei_l, batch_l, full
        torch.cuda.empty_cache()

    # ---------------- device-resident arm: inputs already in HBM -----------------------------------
    ei, batch = spec.build(device)
    feats, x0 = feats_h.to(device), x_h.to(device)
    eng = mod.model.engine_for(ei, feats, batch)
    pred = mod._pred_code() if not ddpm else 1
    sched = list(reversed(range(0, T_STEPS, 1 if ddpm else RATIO)))
    coefs = [mod._step_coef(i, pred) for i in sched]
    xa, xb = x0.clone(), torch.empty_like(x0)
    noise = torch.randn(M, 4, device=device) if ddpm else None
    # small per-GPU batches (the strong-scaled shares: 8 / 4 graphs) are launch-latency bound: the public API's
    # whole-loop CUDA graph (GNN_Diffusion.p_sample_loop_graphed) is what a user would call there
    graphed = bool(w.get("graphed")) or (args.auto_graph and scaling == "strong" and Bl * n <= 8 * 900 and n >= 256)

    def one_step(k, xin, xout):
        c = coefs[k % len(coefs)]
        if ddpm:
            eng.ddpm_step(xin, c, noise if c.t_index != 0 else None, out=xout)
        else:
            eng.ddim_step(xin, c, None, out=xout)

    steps = args.steps
    if graphed:   # whole sampling loops replayed as one CUDA graph each, the remainder of the K steps eagerly
        loops, rem = divmod(args.steps, len(sched))
        for _ in range(2):
            mod.p_sample_loop_graphed((M, 4), feats, ei, batch)
            if rem:
                mod.p_sample_loop_graphed((M, 4), feats, ei, batch, steps_limit=rem)
    for k in range(args.warmup):
        one_step(k, xa, xb); xa, xb = xb, xa
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        clk.begin()
        if world > 1:   # (begin() may have waited for nvidia-smi's first row on some ranks)
            dist.barrier()
            torch.cuda.synchronize()
        ev0.record()
        if graphed:
            for _ in range(loops):
                imgs, _ = mod.p_sample_loop_graphed((M, 4), feats, ei, batch)
            if rem:   # the remaining K mod 30 steps: a partial loop, replayed as its own graph
                imgs, _ = mod.p_sample_loop_graphed((M, 4), feats, ei, batch, steps_limit=rem)
            xa.copy_(imgs[-1])
        else:
            for k in range(steps):
                one_step(k, xa, xb); xa, xb = xb, xa
        if world > 1:  # the ONE collective of the sampling path: gather of the predicted poses
            counts = [(sharding.shard_bounds(w["B"], world, r)[1] - sharding.shard_bounds(w["B"], world, r)[0]) * n
                      if scaling == "strong" else M for r in range(world)]
            sharding.gather_poses(xa, counts)
        ev1.record()
        torch.cuda.synchronize()
        clk.end()
    ms = max_over_ranks(ev0.elapsed_time(ev1), device, world)
    launches = eng.launch_count() - launches0   # (graphed: replays do not pass through the host-side counter; set below)
    ms_per_step = ms / steps
    value = B_global * steps / (ms / 1e3)

    # ---------------- per-kernel times (CUDA events inside the library, same stream) -----------------
    nprof = min(steps, 10)
    if graphed:
        xa, xb = x0.clone(), torch.empty_like(x0)
    eng.set_profiling(True)
    eng.get_profile(reset=True)
    l1 = eng.launch_count()
    for k in range(nprof):
        one_step(k, xa, xb); xa, xb = xb, xa
    torch.cuda.synchronize()
    if graphed:   # the graph replays exactly the launches of the eager steps
        launches = steps * (eng.launch_count() - l1) // nprof
    prof = eng.get_profile(reset=True)
    eng.set_profiling(False)
    stats = eng.graph_stats()
    E, Mt, E_tot = int(ei.shape[1]), eng.num_total, eng.num_edges

    # ---------------- end-to-end arm: public API, pinned host buffers -------------------------------
    e2e = None
    if args.e2e_loops > 0:
        mod.model.invalidate()
        loops = args.e2e_loops if len(sched) <= 30 else max(2, args.e2e_loops // 2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        host_out2 = [torch.empty((len(sched), M, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"next": mod.prefetch(feats_h, spec, None), "k": 0}
        gen = torch.Generator(device=device).manual_seed(7)

        def one_loop():
            k = state["k"]
            # batch i: the whole sampling loop is enqueued (no host sync inside) ...
            ta = time.perf_counter()
            imgs, _ = mod.p_sample_loop((M, 4), *state["next"], generator=gen)
            tb = time.perf_counter()
            # ... and while the GPU samples it, batch i + 1 is uploaded from pinned host memory (features + the graphs'
            # permutations), its edge list written and planned on a side stream into the spare engine
            state["next"] = mod.prefetch(feats_h, spec, None)
            tc = time.perf_counter()
            for s_, img in enumerate(imgs):
                host_out2[k & 1][s_].copy_(img, non_blocking=True)
            if world > 1:
                sharding.gather_poses(imgs[-1], counts)
            done[k & 1].record()
            # the consumer runs one loop behind: batch i - 1's poses are complete on the host before batch i + 1 is
            # enqueued, and the GPU never idles between loops (every byte is still copied inside the timed region)
            if k > 0:
                done[(k - 1) & 1].synchronize()
            state["k"] = k + 1
            state["phases"] = [round(tb - ta, 4), round(tc - tb, 4), round(time.perf_counter() - tc, 4)]

        for _ in range(2):  # untimed warm-up loops (first-use allocations / allocator cache, like the W warm-up steps)
            one_loop()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        loop_s = []
        t0 = time.perf_counter()
        for _ in range(loops):
            tl = time.perf_counter()
            one_loop()
            loop_s.append(time.perf_counter() - tl)
        torch.cuda.synchronize()   # the last batch's poses are on the host
        dt = max_over_ranks(time.perf_counter() - t0, device, world)
        nsteps = loops * len(sched)
        h2d = (feats_h.numel() * 4 + spec.host_bytes()) / len(sched)
        e2e = {"value": B_global * nsteps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(M * 4 * 4),
               "api": "GNN_Diffusion.prefetch + p_sample_loop (%d steps/loop): every loop uploads its pinned host features and the "
                      "graphs' permutations (topology.ExpanderBatchSpec: the edge list is written on the device, scope row N3), "
                      "plans the graph on a side stream into the spare engine (overlapping the previous loop's sampling); every "
                      "step's x_t is read back to pinned host memory, consumed one loop behind; byte counts are per rank" % len(sched),
               "loops": loops, "loop_seconds": [round(x, 4) for x in loop_s],
               "host_phases_s": {"enqueue_loop": state["phases"][0], "prefetch_next": state["phases"][1], "wait_previous": state["phases"][2]}}

    # ---------------- parity of what was just timed: one graph of the workload against the live oracle ----------
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, sample = cpu_baseline(w, steps=args.cpu_steps, keep=True)
        mod.load_state_dict(sample["state"], strict=True)   # the oracle's weights (sharpened attention, see cpu_baseline)
        mod.model.invalidate()
        with torch.no_grad():
            t = torch.full((n,), sample["t"], dtype=torch.long, device=device)
            got, _ = mod.p_sample(sample["x"].to(device), t, sample["t"], cond=None, edge_index=sample["ei"].to(device),
                                  patch_feats=sample["feats"].to(device), batch=sample["batch"].to(device),
                                  noise=sample["noise"].to(device) if sample["noise"] is not None else None)
        want = sample["out"].double()
        parity = float((got.cpu().double() - want).abs().max() / want.abs().max())
        st1 = mod.model._engine.graph_stats()
        parity_note = ("one %d-node graph of the workload (same topology generator, weights, sampler) through the same CUDA path "
                       "(gemm=%s, attn=%s; %d edges on tensor-core tiles, %d on the CSR kernels) vs the CPU oracle's p_sample at "
                       "t=%d: max|y-ref|/max|ref|" % (n, args.gemm, args.attn, st1["dense_edges"], st1["csr_edges"], sample["t"]))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel -----------------------------------------------
    pk = peaks()
    E_dense, E_csr = stats["dense_edges"], stats["csr_edges"]
    hid, D, Hm = 256, 1152, 128
    # rows outside the dense tiles (virtual nodes): on a side stream NEXT TO the per-(tile, head) dense kernel; the persistent
    # hidden-layer kernel (csrc/attn_hidden.cu) owns every SM, so there they run in front of it on the same stream
    try:
        persist_used = mod.model._engine.plan_info().get("persistent_hidden_launches", 0) > 0
    except Exception:
        persist_used = False
    overlapped = {"attn_hidden"} if (stats["dense_edges"] and args.attn == "auto" and not persist_used) else set()
    step_ms_prof = sum(v["ms"] for k, v in prof.items() if k not in overlapped) / nprof
    passes = 3 if args.gemm == "bf16x3" else 1

    def gemm_entry(Mg, N, K, out_bytes_per_elt):
        return {"flops": 2.0 * Mg * N * K, "bytes": 4.0 * Mg * K + 4.0 * N * K + out_bytes_per_elt * Mg * N, "bound": "tensor"}

    # algorithmic work PER LAUNCH of each launch site
    n_extra = 2 * w["V"] * Bl
    work = {
        "prologue": {"flops": 2.0 * M * (64 * Hm + 16 * 4 + 32 * 16), "bytes": 4.0 * M * (Hm + Hm + 8), "bound": "hbm"},
        "mlp2_gemm": gemm_entry(M, D, Hm, 8),
        "qkvs_gemm_first": gemm_entry(Mt, 4 * hid, D, 4),
        "qkvs_gemm_mid": gemm_entry(Mt, 4 * hid, hid, 4),
        "qkvs_gemm_last": gemm_entry(Mt, 4 * D, hid, 4),
        "head_gemm": gemm_entry(M, 32, D, 4),
        # tcgen05 GEMM mode: these tags are the per-layer gather of the promoted extra sources (K and V rows read in
        # fp32 and written as split bf16); exact-fp32 mode: the full image repack
        "pack_hidden": {"flops": 0.0, "bytes": (8.0 * n_extra * 2 * hid) if args.gemm != "fp32" else 4.0 * M * 3 * hid * 2, "bound": "hbm"},
        "pack_last": {"flops": 0.0, "bytes": (8.0 * n_extra * 2 * D) if args.gemm != "fp32" else 4.0 * M * 3 * D * 2, "bound": "hbm"},
        # edge FLOPs of the reference formulation (2C for the score + 2C for the aggregate, per edge and head)
        # HBM bytes: Q / K / V operand images + skip (+ trunk residual on the last layer) in, output planes out
        "attn_dense_hidden": {"flops": 4.0 * E_dense * hid, "bytes": 4.0 * M * hid * 5, "bound": "tensor"},
        "attn_dense_last": {"flops": 4.0 * E_dense * D, "bytes": 4.0 * M * D * 6, "bound": "tensor"},
        # rows served by the CSR kernels: every row without dense tiles, else the rows outside them (Q, skip in, output out per
        # row; K and V rows + the index per in-edge)
        "attn_hidden": {"flops": 4.0 * (E_csr if E_dense else E_tot) * hid,
                        "bytes": 4.0 * ((Mt - M) if E_dense else Mt) * hid * 3 + 4.0 * (E_csr if E_dense else E_tot) * (2 * hid + 1), "bound": "hbm"},
        "attn_last": {"flops": 4.0 * (E_csr if E_dense else E_tot) * D,
                      "bytes": 4.0 * M * D * 5 + 4.0 * (E_csr if E_dense else E_tot) * (2 * D + 1), "bound": "hbm"},
        "head_final": {"flops": 2.0 * M * 32 * 4, "bytes": 4.0 * M * (32 + 12), "bound": "hbm"},
    }
    kernels = {}
    layers_of = {"qkvs_gemm_mid": 2, "attn_dense_hidden": 3, "attn_hidden": 3, "pack_hidden": 3}   # work entries are per layer
    for name, v in prof.items():
        if not v["launches"] or name not in work:
            continue
        w_ = work[name]
        nl = layers_of.get(name, 1)
        ms_step = v["ms"] / nprof
        ms_layer = ms_step / nl          # (attn_hidden: rows outside the dense tiles, on a side stream NEXT TO the dense kernel)
        ent = {"ms_per_launch": round(v["ms"] / v["launches"], 4), "launches_per_step": v["launches"] / nprof,
               "ms_per_step": round(ms_step, 4), "share_of_step": round(ms_step / step_ms_prof, 4), "bound": w_["bound"]}
        if w_["bound"] == "tensor":
            ent["achieved"] = w_["flops"] / (ms_layer * 1e-3) / 1e12
            ent["peak"], ent["unit"] = pk["tensor_sustained"], "TFLOP/s"
        else:
            ent["achieved"] = w_["bytes"] / (ms_layer * 1e-3) / 1e9
            ent["peak"], ent["unit"] = pk["hbm"], "GB/s"
        ent["frac"] = ent["achieved"] / ent["peak"]
        ent["hbm_gbs_algorithmic"] = w_["bytes"] / (ms_layer * 1e-3) / 1e9
        if name in overlapped:
            # rows outside the dense tiles (virtual nodes): launched on a side stream NEXT TO the dense kernel of the same
            # layer, so this span overlaps attn_dense_hidden and is not part of the critical path
            ent["overlapped_with"] = "attn_dense_hidden"
        kernels[name] = ent
    dom_name = max((k for k in kernels if "overlapped_with" not in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dom_name]
    roof = {"bound": dk["bound"], "kernel": dom_name, "achieved": dk["achieved"], "peak": dk["peak"], "unit": dk["unit"],
            "frac": dk["frac"], "traffic": None,
            "peak_source": pk["source"] + (", sustained (kernel timed inside a long step)" if dk["bound"] == "tensor" else ""),
            "share_of_step": dk["share_of_step"], "ms_per_launch": dk["ms_per_launch"], "kernels": kernels}
    if dom_name.startswith("attn_dense"):
        HCd = hid if dom_name.endswith("hidden") else D
        n_scores = 8.0 * Bl * (((n + 127) // 128) * 128) * (((n + 2 * w["V"] + 63) // 64) * 64)  # heads x padded tiles
        roof["note"] = ("achieved = algorithmic edge FLOPs of the reference formulation (4*C per edge and head) / time; the kernel "
                        "executes dense-masked tiles (bitmap density %.2f, padded to 128x64 blocks) with 3 bf16 tensor passes per "
                        "product; the 32-channel layers are paced by the softmax warps (per score ~4 ALU-pipe and 1 SFU instruction on "
                        "16-lane/clk pipes, two warps per scheduler), the 144-channel layer by the tcgen05.mma issue rate (~45 cycles "
                        "per M=128, K=16 instruction for N <= 64: profiles/r2_mma_issue_rate_b200.txt)" % (E_dense / max(1.0, Bl * n * n)))
        roof["executed_tensor_tflops"] = passes * 4.0 * n_scores * (HCd // 8) / (dk["ms_per_launch"] * 1e-3) / 1e12
        roof["scores_per_s"] = n_scores / (dk["ms_per_launch"] * 1e-3)
    elif "gemm" in dom_name:
        roof["note"] = ("algorithmic 2*M*N*K FLOPs; the tensor-core path issues %d bf16 passes per product for fp32 parity, so the "
                        "attainable fraction of the bf16 peak is 1/%d" % (passes, passes))
    # dram__bytes_read + dram__bytes_write of that kernel, per launch, from the ncu capture of this same command
    # (scripts/ncu_traffic.py writes the table from the committed csv; null when the workload was not captured)
    ncu = ROOT / "profiles" / "r2_ncu_traffic.json"
    if ncu.exists():
        try:
            tj = json.loads(ncu.read_text())
            ent = tj.get(args.workload, {}).get(dom_name)
            if ent is not None:
                roof["traffic"] = ent["dram_bytes_per_launch"]
                roof["traffic_source"] = tj.get("_source", "profiles/") + " (ncu --set full, per launch)"
        except Exception:
            pass

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "nodes_per_graph": n, "global_batch_graphs": B_global, "graphs_per_gpu": Bl,
                   "edges_per_gpu": int(E), "edges_per_gpu_with_virtual": int(E_tot),
                   "topology": w["topo"] + (f" {w['degree']}" if w["degree"] else ""),
                   "architecture": w["arch"], "virt_nodes": w["V"],
                   "sampler": "DDPM eps-pred T=300 (300 steps)" if ddpm else "DDIM x0-pred T=300 ratio=10 (30 steps)",
                   "loop": ("device-resident arm: whole 30-step loops replayed as one CUDA graph each (p_sample_loop_graphed), the "
                            "remaining K mod 30 steps as a partial-loop graph; e2e arm: eager p_sample_loop") if graphed else "eager (one fused library call per step)",
                   "gemm_mode": args.gemm, "attn_mode": args.attn, "parallelism": f"graph-shard x{world}",
                   "batch_steps_per_s": steps / (ms / 1e3),
                   "l2_policy": ("per-step working set ~%.2f GB per GPU vs 126 MB L2" % (bytes_per_node() * M / 1e9)) +
                                (" (no flush needed)" if bytes_per_node() * M > 3 * 126e6 else
                                 " (latency regime: the working set is L2-resident BY DESIGN, as it is in the reference's own loop)"),
                   "graph_stats": stats, "workspace_gb": eng.workspace_bytes() / 1e9},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "e2e": e2e,
        "roofline": roof,
        # the three whole-step rates and their fractions of the respective peaks (algorithmic
        # work of the reference formulation / measured step time); the largest is the step's roofline fraction
        "work": step_work(M, E_tot, ms_per_step, pk, passes),
    }
    if parity is not None:
        out["parity_rel_err"] = parity
        out["parity"] = {"rel_err": parity, "tolerance": 1e-4, "what": parity_note}
        out["cpu_baseline"] = cb
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(w, steps=2, keep=False):
    """The oracle (edge-list restatement of the reference formulation) timed on the host cores on a
    bounded sample: ONE graph of the workload per step (B=32 would need >100 GB of per-edge tensors).
    With keep=True also returns the inputs and the output of its first step (the parity probe of run_sample2d)."""
    import oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    ddpm = w["sampler"] == "ddpm"
    ref = oracle.GNNDiffusionRef(steps=T_STEPS, sampling="DDPM" if ddpm else "DDIM", rotation=True, architecture=w["arch"],
                                 virt_nodes=w["V"], model_mean_type=oracle.ModelMeanType.EPSILON if ddpm else oracle.ModelMeanType.START_X,
                                 inference_ratio=1 if ddpm else RATIO).eval()
    # default initialisers give a uniform softmax (every alpha within 1e-5 of 1/deg), which would make the parity probe
    # blind to attention errors: scale the projections so that the attention weights of a row span several decades
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if p.dim() > 1 and "emb" not in name:
                p.mul_(2.25 if ("lin_query" in name or "lin_key" in name) else 1.5)
    n = w["n"]
    if w["topo"] == "dense":
        ei = oracle.dense_edge_index(n)
    else:
        ei = oracle.generate_random_expander(n, w["degree"], rng=np.random.default_rng(graph_seeds(w, 0, 1)[0]),
                                             check_spectral_gap=False).t().contiguous()
    batch = torch.zeros(n, dtype=torch.long)
    g = torch.Generator().manual_seed(0)
    feats, x = torch.randn(n, 1088, generator=g), torch.randn(n, 4, generator=g)
    sched = list(reversed(range(0, T_STEPS, 1 if ddpm else RATIO)))
    sample = None
    with torch.no_grad():
        t = torch.full((n,), sched[0], dtype=torch.long)
        nz = torch.randn(n, 4, generator=g) if ddpm else None
        kw = {"noise": nz} if ddpm else {}
        x1, _ = ref.p_sample(x, t, sched[0], edge_index=ei, patch_feats=feats, batch=batch, **kw)  # warm-up (and parity probe)
        if keep:
            sample = {"x": x, "t": sched[0], "ei": ei, "feats": feats, "batch": batch, "noise": nz, "out": x1, "state": ref.state_dict()}
        x = x1
        t0 = time.perf_counter()
        for k in range(steps):
            i = sched[(k + 1) % len(sched)]
            t = torch.full((n,), i, dtype=torch.long)
            x, _ = ref.p_sample(x, t, i, edge_index=ei, patch_feats=feats, batch=batch)
        dt = time.perf_counter() - t0
    cb = {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{steps} {'DDPM' if ddpm else 'DDIM'} steps of ONE {n}-node graph ({ei.shape[1]} edges) of the workload, torch-cpu "
                    f"edge-list oracle (PyG absent: reference formulation restated), {cores} threads"}
    return (cb, sample) if keep else cb


# ---- configs[3]: ragged 3-D fragment graphs, SE(3) head -------------------------------------------------------------
def c4_inputs(B, seed=3):
    from diffassemble_b200 import topology

    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(2, 21, (B,), generator=g).tolist()
    ei, batch = topology.batch_graphs([topology.dense_edge_index(s) for s in sizes], sizes)
    M = sum(sizes)
    feats = torch.randn(M, 128, generator=g)
    x = torch.cat([torch.tensor([[1.0, 0, 0, 0]]).repeat(M, 1), torch.randn(M, 3, generator=g)], 1)
    return sizes, ei, batch, feats, x


def run_sample3d(args, w):
    import diffassemble_b200 as dab
    import oracle

    world, rank, local_rank, device = init_dist(args)
    B = w["B"]
    sizes, ei_h, batch_h, feats_h, x_h = c4_inputs(B, seed=3 + rank)
    M = sum(sizes)
    torch.manual_seed(0)
    ref = oracle.GNNDiffusion3dRef(steps=T_STEPS, backbone="pointnet", inference_ratio=RATIO, model_mean_type=oracle.ModelMeanType.START_X).eval()
    mod = dab.GNN_Diffusion_3d(steps=T_STEPS, sampling="DDIM", backbone="pointnet", inference_ratio=RATIO,
                               model_mean_type=dab.ModelMeanType.START_X, gemm_mode=args.gemm, attn_mode=args.attn)
    mod.load_state_dict(ref.state_dict(), strict=False)
    mod = mod.to(device)
    ei, batch, feats, x0 = ei_h.to(device), batch_h.to(device), feats_h.to(device), x_h.to(device)
    eng = mod.model.engine_for(ei, feats, batch)
    sched = list(reversed(range(0, T_STEPS, RATIO)))
    coefs = [mod._step_coef(i, mod._pred_code()) for i in sched]
    xa, xb = x0.clone(), torch.empty_like(x0)

    def one_step(k, xin, xout):
        eng.ddim_step(xin, coefs[k % len(coefs)], None, out=xout)

    for k in range(args.warmup):
        one_step(k, xa, xb); xa, xb = xb, xa
    torch.cuda.synchronize()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        ev0.record()
        for k in range(args.steps):
            one_step(k, xa, xb); xa, xb = xb, xa
        ev1.record()
        torch.cuda.synchronize()
        clk.end()
    ms = max_over_ranks(ev0.elapsed_time(ev1), device, world)
    launches = eng.launch_count() - launches0
    # e2e: public API with host buffers, whole 30-step loops
    pin = [t.pin_memory() for t in (feats_h, ei_h, batch_h)]
    host_out = torch.empty((len(sched), M, 7)).pin_memory()
    mod.model.invalidate()

    def one_loop():
        f, e, b = (t.to(device, non_blocking=True) for t in pin)
        imgs, _ = mod.p_sample_loop((M, 7), f, e, b)
        for s_, img in enumerate(imgs):
            host_out[s_].copy_(img, non_blocking=True)
        torch.cuda.synchronize()

    one_loop(); one_loop()
    t0 = time.perf_counter()
    loops = max(2, args.e2e_loops)
    for _ in range(loops):
        one_loop()
    dt = max_over_ranks(time.perf_counter() - t0, device, world)
    # parity + CPU baseline: the oracle on the same batch (small enough to run whole)
    parity, cb = None, None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        with torch.no_grad():
            t = torch.full((M,), sched[0], dtype=torch.long)
            want, _ = ref.p_sample(x_h, t, sched[0], edge_index=ei_h, pcd_feats=feats_h, batch=batch_h)
            t0 = time.perf_counter()
            xx = want
            for k in range(3):
                i = sched[k + 1]
                xx, _ = ref.p_sample(xx, torch.full((M,), i, dtype=torch.long), i, edge_index=ei_h, pcd_feats=feats_h, batch=batch_h)
            dtc = time.perf_counter() - t0
            got, _ = mod.p_sample(x0, t.to(device), sched[0], edge_index=ei, pcd_feats=feats, batch=batch)
        g_, w_ = got.cpu().double(), want.double()
        sign = torch.where((g_[:, :4] * w_[:, :4]).sum(-1, keepdim=True) < 0, -1.0, 1.0)
        g_ = torch.cat([g_[:, :4] * sign, g_[:, 4:]], 1)
        parity = float((g_ - w_).abs().max() / w_.abs().max())
        cb = {"value": B * 3 / dtc, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": f"3 SO(3)/R^3 DDIM steps of the whole batch ({B} graphs, {M} fragments), torch-cpu oracle, {cores} threads"}
    if rank != 0:
        return
    out = {"metric": "denoising-steps/sec on Breaking-Bad-shaped fragment graphs (2-20 nodes, SE(3) head)", "value": B * world * args.steps / (ms / 1e3),
           "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "graphs_per_gpu": B, "nodes_per_gpu": M, "edges_per_gpu": int(ei.shape[1]),
                      "feat_dim": 128, "sampler": "3-D DDIM (R^3 + SO(3)) x0-pred T=300 ratio=10", "gemm_mode": args.gemm, "attn_mode": args.attn,
                      "l2_policy": "latency regime: the whole working set is L2-resident by design"},
           "gpu_launches": int(launches), "clocks": clk.summary(),
           "e2e": {"value": B * world * loops * len(sched) / dt, "unit": UNIT,
                   "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in pin) / len(sched)), "d2h_bytes_per_step": M * 7 * 4,
                   "api": "GNN_Diffusion_3d.p_sample_loop on host tensors, every step's state read back"},
           "roofline": {"bound": "hbm", "kernel": "whole step (launch-latency bound: %d launches of a few microseconds)" % (launches // args.steps),
                        "achieved": 24376.0 * M / (ms / args.steps * 1e-3) / 1e9, "peak": peaks()["hbm"], "unit": "GB/s",
                        "frac": 24376.0 * M / (ms / args.steps * 1e-3) / 1e9 / peaks()["hbm"], "traffic": None,
                        "note": "24 376 compulsory bytes per node (SURVEY.md 8d, D=192); the step is launch-latency bound"}}
    if parity is not None:
        out["parity_rel_err"] = parity
        out["cpu_baseline"] = cb
    print(json.dumps(out))


# ---- configs[4]: one training step ------------------------------------------------------------------------------------
def run_train(args, w):
    import diffassemble_b200 as dab
    import oracle
    from diffassemble_b200 import topology

    world, rank, local_rank, device = init_dist(args)
    import torch.distributed as dist

    B, n = w["B"], w["n"]
    Bl = B // world if args.scaling == "strong" else B
    torch.manual_seed(0)
    mod = dab.GNN_Diffusion(steps=T_STEPS, sampling="DDIM", rotation=True, inference_ratio=RATIO, model_mean_type=dab.ModelMeanType.START_X,
                            gemm_mode=args.gemm, attn_mode=args.attn).to(device)
    ei, batch = topology.DenseBatchSpec(n, Bl).build(device)
    M = Bl * n
    g = torch.Generator().manual_seed(10 + rank)
    feats_h, x0_h = torch.randn(M, 1088, generator=g).pin_memory(), (torch.rand(M, 4, generator=g) * 2 - 1).pin_memory()
    feats, x0 = feats_h.to(device), x0_h.to(device)
    opt = mod.configure_optimizers()
    gen = torch.Generator(device=device).manual_seed(3 + rank)
    params = [p for p in mod.parameters() if p.requires_grad]

    def step(f, x):
        t = torch.randint(0, T_STEPS, (Bl,), device=device, generator=gen)[batch]
        opt.zero_grad(set_to_none=True)
        loss = mod.p_losses(x, t, loss_type="huber", cond=f, edge_index=ei, batch=batch)
        loss.backward()
        if world > 1:
            from diffassemble_b200.training import allreduce_gradients
            allreduce_gradients(params, world)
        opt.step()
        return loss

    for _ in range(args.warmup):
        step(feats, x0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps):
            loss = step(feats, x0)
        ev1.record()
        torch.cuda.synchronize()
    ms = max_over_ranks(ev0.elapsed_time(ev1), device, world)
    # e2e: the batch comes from pinned host memory every step, the loss is read back every step
    t0 = time.perf_counter()
    ne = max(3, args.steps // 2)
    for _ in range(ne):
        l_ = step(feats_h.to(device, non_blocking=True), x0_h.to(device, non_blocking=True)).item()
    dt = max_over_ranks(time.perf_counter() - t0, device, world)
    cb = None
    if rank == 0 and not args.no_cpu_baseline:
        from transformers.optimization import Adafactor

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Bc = 4
        ref = oracle.GNNDiffusionRef(steps=T_STEPS, sampling="DDIM", rotation=True, inference_ratio=RATIO, model_mean_type=oracle.ModelMeanType.START_X)
        opt_c = Adafactor(ref.parameters())
        eic, batchc = oracle.batch_graphs([oracle.dense_edge_index(n)] * Bc, [n] * Bc)
        fc, xc = torch.randn(Bc * n, 1088), torch.rand(Bc * n, 4)

        def cstep():
            t = torch.randint(0, T_STEPS, (Bc,))[batchc]
            opt_c.zero_grad()
            l = ref.p_losses(xc, t, loss_type="huber", edge_index=eic, patch_feats=fc, batch=batchc)
            l.backward()
            opt_c.step()

        cstep()
        tc = time.perf_counter()
        for _ in range(3):
            cstep()
        dtc = (time.perf_counter() - tc) / 3
        cb = {"value": Bc / dtc, "unit": "graphs/s", "cores": cores, "kind": "port",
              "sample": f"3 training steps (oracle autograd + transformers Adafactor) on {Bc} 144-node graphs, {cores} threads"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    Bg = Bl * world
    f_step = 3.0 * (M * flops_per_node() + int(ei.shape[1]) * flops_per_edge())   # forward + data gradient + weight gradient
    out = {"metric": "training graphs/sec (forward + backward + Adafactor) on 12x12 puzzles", "value": Bg * args.steps / (ms / 1e3), "unit": "graphs/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32 (3-pass split-bf16 tensor-core GEMMs)", "data": "synthetic",
           "config": {"workload": args.workload, "graphs_per_gpu": Bl, "nodes_per_graph": n, "loss": "huber", "optimizer": "Adafactor (fused)",
                      "gemm_mode": args.gemm, "attn_mode": args.attn, "final_loss": float(loss.item()),
                      "l2_policy": "per-step working set ~%.2f GB vs 126 MB L2" % (3 * bytes_per_node() * M / 1e9)},
           "gpu_launches": None, "clocks": clk.summary(),
           "e2e": {"value": Bg * ne / dt, "unit": "graphs/s", "h2d_bytes_per_step": int(feats_h.numel() * 4 + x0_h.numel() * 4), "d2h_bytes_per_step": 4,
                   "api": "GNN_Diffusion.p_losses + backward + FusedAdafactor.step; batch uploaded from pinned host memory and the loss read back every step"},
           "roofline": {"bound": "tensor", "kernel": "whole training step", "achieved": f_step / (ms / args.steps * 1e-3) / 1e12,
                        "peak": peaks()["tensor_sustained"], "unit": "TFLOP/s", "frac": f_step / (ms / args.steps * 1e-3) / 1e12 / peaks()["tensor_sustained"],
                        "traffic": None, "note": "algorithmic FLOPs: 3 x forward (forward, data gradient, weight gradient)"}}
    if cb is not None:
        out["cpu_baseline"] = cb
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    if w["kind"] != "sample2d":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm is defined for the 2-D sampling workloads (c2 / c3)"}))
        return
    steps = max(1, min(args.steps, 3))
    cb = cpu_baseline(w, steps=steps)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
           "steps": steps, "warmup": 1, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "nodes_per_graph": w["n"], "graphs_per_step": 1,
                      "note": "bounded sample: one graph per step on the host cores"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_exphander60_v8", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the global batch of 32 graphs is sharded over the ranks (BASELINE configs[2]); weak: 32 graphs per GPU")
    ap.add_argument("--gemm", default="bf16x3", choices=["fp32", "bf16x3"])
    ap.add_argument("--attn", default="auto", choices=["csr", "auto"])
    ap.add_argument("--e2e-loops", type=int, default=10)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-auto-graph", dest="auto_graph", action="store_false",
                    help="keep the eager loop even for small strong-scaled per-GPU batches")
    ap.add_argument("--graphed", action="store_true", help="replay the whole sampling loop as one CUDA graph (p_sample_loop_graphed)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="override the workload's global batch (development: e.g. 4 = the per-GPU share of the 8-GPU strong-scaling run)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    w = dict(WORKLOADS[args.workload])
    if args.global_batch > 0:
        w["B"] = args.global_batch
    if args.graphed:
        w["graphed"] = True
    {"sample2d": run_sample2d, "sample3d": run_sample3d, "train": run_train}[w["kind"]](args, w)


if __name__ == "__main__":
    main()
