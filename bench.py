#!/usr/bin/env python
"""Benchmark of the multi-particle stochastic-node hot path (BASELINE.json metric):
particle-samples/s for the IWAE / VIMCO objective forward+backward at K=50, plus % of HBM peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload iwae|vimco] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input of config 2
(K=50 particles, B=1024 batch columns per GPU, Z=40 latent, X=784 observed; fp32):
    sample z ~ q (Philox in-kernel) -> log q(z|x) -> log p(z) -> fused Bernoulli-likelihood +
    importance-weighted objective forward+backward (dprobs, dlogp, dlogq) -> backward of the two
    Normal log-densities -> pathwise backward of the sample (dmean, dstd)
with the leaves at the path boundary of SURVEY.md §8(d): mean/std [B,Z], probs [K,B,X] (the decoder
output), x [B,X] and the decoder's upstream gradient dz [K,B,Z].  The MLP GEMMs are not part of the path.

  value : device-resident inputs, C-ABI calls (through the ctypes binding), CUDA events.
  e2e   : the same step through the public Python API with PINNED HOST tensors as inputs
          (host->device and device->host copies inside the timed region).
  cpu_baseline / --impl reference : the CPU oracle port (C, OpenMP, all host threads) of the same step
          on a bounded sample (the reference is pure Python and cannot travel to the GPU box).
Rank 0 prints ONE JSON line.  N>1: one process per GPU (torchrun), batch columns sharded (weak
scaling), no data-path collective; the scalar objective is summed in place by the fused launch and all-reduced once
per 128 steps (the loss-reporting interval).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "zhusuan-pytorch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

K_PART, B_COLS, Z_DIM, X_DIM = 50, 1024, 40, 784
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--workload", choices=["iwae", "vimco"], default="iwae")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="replay the step from a CUDA graph (1) or launch eagerly (0)")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_port_inputs(B, vimco, seed=0):
    import numpy as np
    rng = np.random.RandomState(seed)
    K, Z, X = K_PART, Z_DIM, X_DIM
    probs = (1.0 / (1.0 + np.exp(-2.0 * rng.standard_normal((K, B, X))))).astype(np.float32)
    x = (rng.uniform(size=(B, X)) < 0.5).astype(np.float32)
    if vimco:
        pq = (1.0 / (1.0 + np.exp(-rng.standard_normal((B, Z))))).astype(np.float32)
        u = rng.uniform(size=(K, B * Z)).astype(np.float32)
        return dict(probs=probs, x=x, pq=pq, u=u)
    mean = (0.5 * rng.standard_normal((B, Z))).astype(np.float32)
    std = np.exp(0.3 * rng.standard_normal((B, Z))).astype(np.float32)
    eps = rng.standard_normal((K, B * Z)).astype(np.float32)
    dz_up = (1e-3 * rng.standard_normal((K, B, Z))).astype(np.float32)
    return dict(probs=probs, x=x, mean=mean, std=std, eps=eps, dz_up=dz_up)


def cpu_port_step(O, inp, vimco):
    """The hot path on the CPU oracle (C + OpenMP): same stages as the GPU step."""
    import numpy as np
    K, Z, X = K_PART, Z_DIM, X_DIM
    B = inp["x"].shape[0]
    if vimco:
        pq = inp["pq"]
        z = O.bernoulli_sample(pq, inp["u"], K, B * Z).reshape(K, B, Z)
        logq = O.bernoulli_logpmf_fwd(z, pq, K, B, Z)
        logpz = O.bernoulli_logpmf_fwd(z, np.full((B, Z), 0.5, np.float32), K, B, Z)
        r = O.iw_bernoulli_step(O.VIMCO, inp["probs"], inp["x"], logpz, logq)
        dpq = O.bernoulli_logpmf_bwd(r["dlogq"], z, pq, K, B, Z)
        return r["cost"].mean(), r["dprobs"], dpq
    mean, std, eps = inp["mean"], inp["std"], inp["eps"]
    z = O.normal_sample(mean, std, eps, K, B * Z).reshape(K, B, Z)
    logq = O.normal_logprob_fwd(z, mean, std, K, B, Z)
    zeros, ones = np.zeros((B, Z), np.float32), np.ones((B, Z), np.float32)
    logpz = O.normal_logprob_fwd(z, zeros, ones, K, B, Z)
    r = O.iw_bernoulli_step(O.SGVB, inp["probs"], inp["x"], logpz, logq)
    dz_p, _, _ = O.normal_logprob_bwd(r["dlogp"], z, zeros, ones, K, B, Z)
    dz_q, dm_q, ds_q = O.normal_logprob_bwd(r["dlogq"], z, mean, std, K, B, Z)
    dz = (inp["dz_up"] + dz_p + dz_q).reshape(K, B * Z)
    dm_s, ds_s = O.normal_sample_bwd(dz, eps, mean, std, K, B * Z)
    return r["cost"].mean(), r["dprobs"], dm_q + dm_s, ds_q + ds_s


def time_cpu_port(vimco, B_sample, reps, warm):
    from oracle import zs_oracle as O
    O.build()
    cores = O.set_threads(os.cpu_count() or 1)
    inp = cpu_port_inputs(B_sample, vimco)
    for _ in range(warm):
        cpu_port_step(O, inp, vimco)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_port_step(O, inp, vimco)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return dict(value=K_PART * B_sample / med, unit="particle-samples/s", cores=cores, kind="port",
                sample="oracle C port (OpenMP, %d threads) of the same step on %d of the %d batch columns, "
                       "median of %d runs, %.1f ms each" % (cores, B_sample, B_COLS, reps, med * 1e3)), med


def run_reference_arm(args):
    """`--impl reference`: the reference is pure Python (torch eager) and does not exist on the GPU
    box, so the CPU arm is the oracle port of its algorithm on all host threads (kind: "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vimco = args.workload == "vimco"
    B_sample = 256
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 3))
    cb, med = time_cpu_port(vimco, B_sample, steps, warm)
    line = {
        "impl": "reference", "metric": "particle-samples/sec for IWAE/VIMCO fwd+bwd (K=50)", "value": cb["value"],
        "unit": "particle-samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(vimco), "K": K_PART, "B_per_step": B_sample, "Z": Z_DIM, "X": X_DIM},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "particle-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(vimco):
    return ("vimco_bernoulli_latents" if vimco else "iwae_normal_latents") + "_K50_B1024_Z40_X784_path"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [c.strip() for c in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- GPU arm
class PathStep(object):
    """Device-resident hot-path step through the C-ABI binding (zhusuan._backend)."""

    def __init__(self, torch, be, vimco, device, seed):
        self.torch, self.be, self.vimco, self.dev = torch, be, vimco, device
        K, B, Z, X = K_PART, B_COLS, Z_DIM, X_DIM
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        rn = lambda *s: torch.randn(*s, device=device, generator=g)
        self.probs = torch.sigmoid(2.0 * rn(K, B, X)).contiguous()
        self.x = (torch.rand(B, X, device=device, generator=g) < 0.5).float()
        self.offset = 0
        self.seed = seed
        if vimco:
            self.pq = torch.sigmoid(rn(B, Z)).contiguous()
            self.prior = torch.full((B, Z), 0.5, device=device)
        else:
            self.mean = (0.5 * rn(B, Z)).contiguous()
            self.std = torch.exp(0.3 * rn(B, Z)).contiguous()
            self.zeros = torch.zeros(B, Z, device=device)
            self.ones = torch.ones(B, Z, device=device)
            self.dz_up = (1e-3 * rn(K, B, Z)).contiguous()
        self.launches_per_step = 0
        self.out = None
        # data-parallel runs: the per-column objectives of successive steps are summed in place by the fused launch
        # (zs_iw_bernoulli_fused_accumulate) and reduced once per bucket of steps, not once per step
        self.cost_sum = None

    def algorithmic_bytes(self):
        """Compulsory HBM bytes of one step with the fused design."""
        K, B, Z, X = K_PART, B_COLS, Z_DIM, X_DIM
        kbx, kbz, kb, bz, bx = 4 * K * B * X, 4 * K * B * Z, 4 * K * B, 4 * B * Z, 4 * B * X
        fused = 2 * kbx + bx + 4 * kb + 4 * B          # probs R, dprobs W, x R, other/logq R, dlogp/dlogq W, cost W
        if self.vimco:
            small = (bz + kbz + 2 * kb) + (kb + kbz + 2 * bz)            # latent fwd (W z, logq, logp), latent bwd
        else:
            small = (2 * bz + kbz + 2 * kb) + (2 * kb + 2 * kbz + 4 * bz)  # fwd: R mean,std W z,logq,logp; bwd: R g's,z,dz_up
        return fused, fused + small

    def step(self):
        """latent forward (sample + log q + log p(z)) -> fused likelihood+objective fwd+bwd -> latent
        backward: three launches of our kernels per step."""
        be = self.be
        K, B, Z, X = K_PART, B_COLS, Z_DIM, X_DIM
        n0 = be.launch_count
        self.offset += 4
        if self.vimco:
            z, logq, logpz = be.bernoulli_latent_fwd(self.pq, be.KBCAST, K, B, Z, seed=self.seed, offset=self.offset)
            r = be.iw_bernoulli_fused(be.VIMCO, self.probs, self.x, logpz, logq, 1.0 / B, **self._acc())
            dpq = be.bernoulli_latent_bwd(r["dlogq"], z, self.pq, be.KBCAST, K, B, Z)
            self.out = (r["cost"], r["dprobs"], dpq)
        else:
            z, logq, logpz = be.normal_latent_fwd(self.mean, self.std, be.KBCAST, K, B, Z, seed=self.seed,
                                                  offset=self.offset)
            r = be.iw_bernoulli_fused(be.SGVB, self.probs, self.x, logpz, logq, 1.0 / B, **self._acc())
            dm, ds = be.normal_latent_bwd(r["dlogq"], r["dlogp"], self.dz_up, z, self.mean, self.std, be.KBCAST, K, B,
                                          Z, reparameterized=True)
            self.out = (r["cost"], r["dprobs"], dm, ds)
        self.launches_per_step = be.launch_count - n0
        return self.out

    def _acc(self):
        return {} if self.cost_sum is None else dict(out={"cost": self.cost_sum}, accumulate_cost=True)

    def fused_only(self, other, logq, out=None):
        return self.be.iw_bernoulli_fused(self.be.VIMCO if self.vimco else self.be.SGVB, self.probs, self.x, other,
                                          logq, 1.0 / B_COLS, out=out)


def api_step_host(torch, zs, vimco, host):
    """The same step through the PUBLIC Python API with pinned HOST tensors as leaves (the e2e
    measurement): the package uploads what the kernels need, runs them, and returns the loss and the
    gradients in host memory.  For the host-resident likelihood tensor the objective takes the
    chunk-pipelined route (zs_iw_step_host)."""
    from zhusuan.distributions import Bernoulli, Normal
    from zhusuan.framework import BayesianNet
    from zhusuan.variational import ImportanceWeightedObjective
    K = K_PART
    cpu = torch.device("cpu")
    probs = host["probs"].detach().requires_grad_()
    x = host["x"]

    class Gen(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            if vimco:
                self.bernoulli("z", probs=host["prior"], n_samples=K, reduce_sum_dims=[2])
            else:
                self.normal("z", mean=host["zeros"], std=host["ones"], is_reparameterized=False, n_samples=K,
                            reduce_sum_dims=[2])
            self.sn(Bernoulli(probs=probs), name="x", reduce_sum_dims=[2])
            return self

    class Var(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            if vimco:
                self.sn(Bernoulli(probs=a), name="z", n_samples=K, reduce_sum_dims=[2])
            else:
                self.sn(Normal(mean=a, std=b), name="z", n_samples=K, reduce_sum_dims=[2])
            return self

    if vimco:
        a, b = host["pq"].detach().requires_grad_(), None
    else:
        a, b = host["mean"].detach().requires_grad_(), host["std"].detach().requires_grad_()
    obj = ImportanceWeightedObjective(Gen(device=cpu), Var(device=cpu), axis=0, estimator="vimco" if vimco else "sgvb")
    loss = obj({"x": x})
    loss.backward()
    assert probs.grad is not None and probs.grad.device.type == "cpu" and a.grad is not None
    return float(loss)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from zhusuan import _backend as be
    import zhusuan as zs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    be.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    be.load()
    vimco = args.workload == "vimco"
    pg = world > 1 or os.environ.get("ZS_BENCH_FORCE_PG") == "1"  # the env knob: a 1-rank NCCL group, for diagnosis

    def init_pg():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("nccl", device_id=dev, rank=0, world_size=1)
            dist.all_reduce(torch.zeros(1, device=dev))

    # ZS_BENCH_PG_FIRST=1 restores the old order (communicator first, then the step's buffers): measured on one GPU,
    # a step whose 160 MB tensors are allocated AFTER NCCL's own device allocations runs ~5 us slower
    pg_first = os.environ.get("ZS_BENCH_PG_FIRST") == "1"
    if pg and pg_first:
        init_pg()
    pad_mb = float(os.environ.get("ZS_BENCH_PAD_MB", "0"))  # dev knob: shift the placement of the step's tensors
    pad = torch.empty(int(pad_mb * (1 << 20)), dtype=torch.uint8, device=dev) if pad_mb > 0 else None
    ps = PathStep(torch, be, vimco, dev, seed=1234 + rank)
    if world > 1 or os.environ.get("ZS_BENCH_ACCUM") == "1":  # the env knob isolates the cost of the running sum
        ps.cost_sum = torch.zeros(B_COLS, device=dev)
    for _ in range(3):  # sizes the caching allocator before anything else allocates on the device
        ps.step()
    torch.cuda.synchronize()
    if pg and not pg_first:
        init_pg()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- warm-up (also sizes the caching allocator) and optional CUDA-graph capture of the step
    for _ in range(3):
        ps.step()
    torch.cuda.synchronize()
    graph = None
    if args.graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ps.step()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    ps.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        except Exception as e:  # capture unsupported: fall back to eager launches, and say so
            graph = None
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (e,))

    # The only cross-rank exchange of this path is the scalar objective.  The fused launch adds each step's
    # per-column objectives into ps.cost_sum; once per LOSS_BUCKET steps (the loss-reporting interval) the sum is
    # reduced and all-reduced.  The collective is ordered IN the compute stream on purpose: measured at N=2, an
    # NCCL kernel that overlaps the step (side stream) holds SMs while it waits for its peer, the persistent
    # 148-CTA likelihood kernel then runs one CTA short and takes a second pass -- 11 % slower steps.  In-stream,
    # the cost is ~30 us per bucket.
    LOSS_BUCKET = int(os.environ.get("ZS_BENCH_BUCKET", "128"))
    reduced = torch.zeros(1, device=dev)
    state = {"i": 0}

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            ps.step()
        if world > 1:
            state["i"] += 1
            if state["i"] % LOSS_BUCKET == 0 and os.environ.get("ZS_BENCH_NOBUCKET") != "1":
                torch.sum(ps.cost_sum, dim=0, keepdim=True, out=reduced)
                ps.cost_sum.zero_()
                reduced.mul_(1.0 / (LOSS_BUCKET * B_COLS * world))
                dist.all_reduce(reduced, op=dist.ReduceOp.SUM)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    # clocks: the timed region can be shorter than nvidia-smi's sampling period; keep the same step
    # running (untimed) until a few samples under load exist
    t_clock_end = t_end
    if rank == 0:
        while len([1 for t, _ in clocks.rows if t >= t_start]) < 6 and time.perf_counter() - t_end < 3.0:
            for _ in range(200):  # compute only: no collective here, the other ranks are not in this loop
                if graph is not None:
                    graph.replay()
                else:
                    ps.step()
            torch.cuda.synchronize()
            t_clock_end = time.perf_counter()
        clocks.stop()
    ms_by_rank = [ms]
    if world > 1:
        allms = [torch.zeros(1, device=dev) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([ms], device=dev))
        ms_by_rank = [float(t) for t in allms]
        ms = max(ms_by_rank)
    ms_per_step = ms / args.steps
    value = world * K_PART * B_COLS / (ms_per_step * 1e-3)

    # --- dominant kernel alone (roofline): CUDA events on the launching stream
    fused_bytes, step_bytes = ps.algorithmic_bytes()
    other = torch.randn(K_PART, B_COLS, device=dev) - 55.0
    logq = torch.randn(K_PART, B_COLS, device=dev) + 30.0
    outbuf = ps.fused_only(other, logq)  # output buffers reused by every timed launch
    for _ in range(10):
        ps.fused_only(other, logq, out=outbuf)
    torch.cuda.synchronize()
    # launched from a CUDA graph (20 launches per replay) so that the host's launch rate -- a ctypes call plus the
    # binding's bookkeeping costs about as much as the kernel runs -- is not what the events measure
    per_graph, kgraph = 20, None
    if args.graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                kgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(kgraph, stream=side):
                    for _ in range(per_graph):
                        ps.fused_only(other, logq, out=outbuf)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        except Exception as e:
            kgraph = None
            sys.stderr.write("CUDA graph capture of the fused kernel failed, timing eager launches: %r\n" % (e,))
    reps = 200
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if kgraph is not None:
        kgraph.replay()
        torch.cuda.synchronize()
        k0.record()
        for _ in range(reps // per_graph):
            kgraph.replay()
        k1.record()
    else:
        k0.record()
        for _ in range(reps):
            ps.fused_only(other, logq, out=outbuf)
        k1.record()
    torch.cuda.synchronize()
    fused_ms = k0.elapsed_time(k1) / reps
    peak, peak_kind = hbm_peak()
    achieved = fused_bytes / (fused_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch_" + args.workload)
        except Exception:
            traffic = None

    line = {
        "metric": "particle-samples/sec for IWAE/VIMCO fwd+bwd (K=50)", "value": value, "unit": "particle-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(vimco), "K": K_PART, "B_per_gpu": B_COLS, "Z": Z_DIM, "X": X_DIM,
                   "estimator": "vimco" if vimco else "sgvb", "parallelism": "batch columns sharded x%d" % world,
                   "l2": "inputs+outputs of a step (%.0f MB) exceed the 126 MB L2; no explicit flush" % (step_bytes / 1e6),
                   "launch": "cuda-graph replay" if graph is not None else "eager launches",
                   "timed_wall_s": t_end - t_start,
                   "ms_per_step_by_rank": [round(v / args.steps, 6) for v in ms_by_rank]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_iw_bernoulli_boxf (zs_iw_bernoulli_fused)", "kernel_ms": fused_ms,
                     "algorithmic_bytes_per_launch": fused_bytes, "peak_kind": peak_kind + " (MEASURED_PEAKS.json)"
                     if peak_kind == "measured" else "fallback (B200_PROFILING.md)",
                     "step_algorithmic_bytes": step_bytes,
                     "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                     "survey_bytes_per_particle_sample": 10074 if not vimco else 9911,
                     "survey_equiv_frac": (10074 if not vimco else 9911) * K_PART * B_COLS / (ms_per_step * 1e-3) / 1e9 / peak},
        "gpu_launches": ps.launches_per_step * args.steps,
        "launches_per_step": ps.launches_per_step,
    }

    if rank == 0:
        line["clocks"] = clocks.summary(t_start, t_clock_end)

    # --- e2e through the public API with pinned host buffers
    if not args.no_e2e:
        pin = lambda t: t.detach().cpu().pin_memory()
        host = {"probs": pin(ps.probs), "x": pin(ps.x)}
        if vimco:
            host.update(pq=pin(ps.pq), prior=pin(ps.prior))
        else:
            host.update(mean=pin(ps.mean), std=pin(ps.std), zeros=pin(ps.zeros), ones=pin(ps.ones))
        n_e2e = max(3, args.e2e_steps)
        for _ in range(3):
            api_step_host(torch, zs, vimco, host)
        barrier()
        n0 = be.launch_count
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            api_step_host(torch, zs, vimco, host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt)
        kbx, bx, kbz, kb, bz = (4 * K_PART * B_COLS * X_DIM, 4 * B_COLS * X_DIM, 4 * K_PART * B_COLS * Z_DIM,
                                4 * K_PART * B_COLS, 4 * B_COLS * Z_DIM)
        n_par = 1 if vimco else 2
        # uploads: probs, x, variational + prior parameters (twice: the reference protocol reads .tensor twice per
        # step); downloads: dprobs, the two z draws the API returns to its host-resident caller, the per-column
        # costs, parameter gradients.  The [K,B] log-probabilities and their gradients stay on the device.
        h2d = kbx + bx + 2 * 2 * n_par * bz
        d2h = kbx + 2 * kbz + 4 * B_COLS + n_par * bz
        line["e2e"] = {"value": world * K_PART * B_COLS / dt, "unit": "particle-samples/s",
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
                       "api": "zhusuan.variational.ImportanceWeightedObjective(...)({'x': x}); loss.backward(), every "
                              "leaf (probs, x, parameters) and every result (loss, gradients) in host memory",
                       "launches_per_step": (be.launch_count - n0) // n_e2e,
                       "pcie_floor_ms": "3.3-3.5 (H2D and D2H of the 160.6 MB likelihood tensor / gradient at once, "
                                        "measured 93-98 GB/s aggregate)"}

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cb, _ = time_cpu_port(vimco, 128, 12, 2)
        line["cpu_baseline"] = cb
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if pg:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
