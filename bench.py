#!/usr/bin/env python
"""bench.py -- pseudo-labelled samples/s of the UBPL hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU baseline (oracle port)

One step = one pass of the fused chain over one synthetic batch resident in HBM:
K1 back-warp+flip+arg-max decode of all M*K teacher views -> K2 dispersion + selection ->
K3 Gaussian render + masked joint-MSE forward + gradient -> K4 mean-teacher EMA of an HG2-sized
parameter set.  Weak scaling: every rank owns a full per-GPU batch (no data-path collective; the
global-quantile configs all-reduce the selection histograms over NCCL).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

# BASELINE.json configs (per-GPU batch under weak scaling)
CONFIGS = {
    "c1": dict(B=16, K=4, M=1, S=2, J=14, H=64, W=64, select="fixed", hg="hg2_j14",
               desc="MT_UBPL LSP J=14 B=16 K=4 (the reference's CPU-runnable case)"),
    "c2": dict(B=256, K=8, M=1, S=2, J=14, H=64, W=64, select="fixed", hg="hg2_j14",
               desc="LSP J=14 B=256 K=8 mean-teacher pseudo-labels + EMA"),
    "c3": dict(B=128, K=8, M=2, S=2, J=9, H=64, W=64, select="fixed", hg="hg2_j9",
               desc="DualPose_UBPL FLIC J=9 B=1024/8 per GPU K=8 dual teachers"),
    "c4": dict(B=256, K=16, M=1, S=2, J=17, H=64, W=64, select="quantile", hg="hg2_j17",
               desc="AP-10K J=17 B=2048/8 per GPU K=16 global-quantile threshold (NCCL histogram all-reduce)"),
    "c5": dict(B=256, K=16, M=1, S=2, J=32, H=128, W=128, select="fixed", hg="hg2_j32",
               desc="fly J=32 128x128 K=16 bandwidth stress, one resident chunk of 256 of the 4096 samples per step "
                    "(9.6 GB; pipeline.stream_chunks streams the rest)"),
}
DIST_THR_MAX = 3.0          # no reference default exists; ~half of the joints pass on the synthetic data
METRIC = "pseudo-labelled samples/sec"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE warp_decode_kernel launch, from the ncu --set full captures
# profiles/r02/final_k1_k3_full.ncu-rep (c2, K1 with the EMA in its tail: 539.10 MB read = 469.76 MB of teacher maps +
# 67.4 MB of EMA operands, 4.74 MB written before the launch ends -- the EMA's 33.7 MB of results are still in L2 then)
# and, for c4, round 1's profiles/r01f_c4_k1_select_k3_full.ncu-rep (1141.15 MB + 4.64 MB vs 1140.85 MB algorithmic)
NCU_TRAFFIC = {"c2": 543.84e6, "c4": 1145.79e6}
# c4 at N = 1 (this repo, B200, `python bench.py --config c4`, profiles/r02/): the denominator of the collective
# block's `vs_n1` when the driver's N > 1 runs time c4 beside the headline config
C4_N1 = {"value": 885398.0, "ms_per_step": 0.2891, "source": "profiles/r02/final_bench_c4.json (builder-run, 1 B200); the N = 1 run of this "
         "script carries its own `collective` block, which is the denominator to use"}


def algorithmic_bytes_per_sample(c):
    """teacher maps read once + student read once + gradient written once + target written once."""
    return 4 * c["H"] * c["W"] * c["J"] * (c["M"] * c["K"] + c["S"] + c["S"] + 1)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while not self.lines and time.time() - t0 < timeout and self.proc is not None:
            time.sleep(0.02)

    def stop(self, t_from=0.0, t_to=float("inf")):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < t_from or ts > t_to:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm.  kind "reference": the reference's OWN functions (utils/augment.py, process.py, evaluation.py,
# business.py, losses.py, parameters.py), imported unmodified by oracle/ref_import.py from /root/reference or,
# on the GPU box, from the byte-for-byte staged copy under baseline/_ref (tools/stage_reference.py), composed
# by oracle/ref_chain.py.  kind "port": the numpy oracle, only when no reference tree is present.  All host
# cores, one single-threaded process per core, a bounded sample of the workload per step.
# ---------------------------------------------------------------------------------------------------
def cpu_kind():
    import ref_import
    return "reference" if ref_import.reference_available() else "port"


def _cpu_worker(args):
    cfgname, nb, seed, kind = args
    import torch
    torch.set_num_threads(1)
    import ubpl_b200  # noqa: F401
    from ubpl_b200 import synth
    c = CONFIGS[cfgname]
    d = synth.make_batch(B=nb, K=c["K"], J=c["J"], H=c["H"], W=c["W"], M=c["M"], S=c["S"], seed=seed)
    if kind == "reference":
        import ref_import
        import ref_chain
        ref = ref_import.load_reference()
        t0 = time.perf_counter()
        ref_chain.reference_chain(ref, d, select=c["select"], distThrMax=DIST_THR_MAX)
        return time.perf_counter() - t0
    import ubpl_oracle as O
    n = {k: v.numpy() for k, v in d.items()}
    t0 = time.perf_counter()
    O.pseudo_label_chain(n["teacher"], n["student"], n["theta"], n["flip"], n["center"], n["scale"], n["islabeled"],
                         select=c["select"], distThrMax=DIST_THR_MAX)
    return time.perf_counter() - t0


class CpuChain:
    """A pool of `procs` single-threaded workers; step() runs `per_proc` samples of the config's chain in each and
    returns (samples, seconds) by the wall clock of the pool."""

    def __init__(self, cfgname, procs, kind):
        import multiprocessing as mp
        self.cfgname, self.procs, self.kind = cfgname, procs, kind
        self.pool = mp.get_context("spawn").Pool(procs)
        self.pool.map(_cpu_worker, [(cfgname, 1, 7 + i, kind) for i in range(procs)])     # imports, first-call costs
        self.calls = 0

    def step(self, per_proc):
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(self.cfgname, per_proc, 1388 + 1000 * self.calls + i, self.kind) for i in range(self.procs)])
        self.calls += 1
        return self.procs * per_proc, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_ema_seconds(c, kind):
    """One update_ema_variables over an HG2 parameter set: the reference's own function on two reference
    StackedHourglass models (kind "reference", utils/parameters.py:4-8, all torch threads), else numpy."""
    if kind == "reference":
        import torch
        import ref_import
        import ref_chain
        ref = ref_import.load_reference()
        model, ema = ref_chain.make_hourglass_pair(ref, c["J"])
        ref_chain.reference_ema(ref, model, ema)                 # warm
        t0 = time.perf_counter()
        for _ in range(3):
            ref_chain.reference_ema(ref, model, ema)
        return (time.perf_counter() - t0) / 3
    import numpy as np
    shapes = json.load(open(os.path.join(ROOT, "ubpl-poseestimation_b200", "hg_param_shapes.json")))[c["hg"]]
    rng = np.random.default_rng(0)
    ps = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    es = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    a = np.float32(0.75)
    t0 = time.perf_counter()
    for e, p in zip(es, ps):
        e *= a
        e += p * (np.float32(1) - a)
    return time.perf_counter() - t0


def cpu_sample_note(kind, n, procs, per_proc, cfgname):
    src = ("the reference's own functions (oracle/ref_chain.py over the unmodified utils/*.py) + update_ema_variables on "
           "reference HG2 models" if kind == "reference" else "oracle/ubpl_oracle.py chain + numpy EMA")
    return "%d samples of %s per step (%d single-threaded processes x %d): %s, EMA pro-rated per batch" % (
        n, cfgname, procs, per_proc, src)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                     # rank 0 alone runs the CPU arm
    c = CONFIGS[args.config]
    kind = cpu_kind()
    procs = max(1, min(os.cpu_count() or 1, 64))
    # every step is a bounded sample (procs x per_proc samples), sized so that the driver's --steps/--warmup end
    # within a few minutes: ~25 samples/s per core at c2 shapes, ~1 sample/s per core at c5's
    heavy = c["H"] * c["W"] * c["J"] * c["K"] * c["M"] > 64 * 64 * 14 * 8 * 4
    per_proc = max(1, int(args.ref_samples_per_proc)) if args.ref_samples_per_proc else (1 if heavy else 4)
    chain = CpuChain(args.config, procs, kind)
    ema_s = cpu_ema_seconds(c, kind)
    for _ in range(args.warmup):
        chain.step(1)
    vals, t_steps = [], []
    for _ in range(args.steps):
        n, dt = chain.step(per_proc)
        tot = dt + ema_s * n / c["B"]               # one EMA per full batch of B samples, pro-rated
        vals.append(n / tot)
        t_steps.append(tot)
    chain.close()
    v = sum(t_steps) and (procs * per_proc * len(t_steps)) / sum(t_steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_steps) / len(t_steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": procs, "kind": kind,
                         "sample": cpu_sample_note(kind, procs * per_proc, procs, per_proc, args.config)},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cfgname):
    """The `config` keys both arms share (the GPU arm adds measurement details under other names)."""
    c = CONFIGS[cfgname]
    return {"workload": cfgname + ": " + c["desc"], "per_gpu_batch": c["B"], "K": c["K"], "M": c["M"], "J": c["J"],
            "S": c["S"], "heatmap": [c["H"], c["W"]], "select": c["select"], "distThrMax": DIST_THR_MAX}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def _timed_replays(step_fn, n, barrier, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(n):
        step_fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def measure_config(cfgname, args, env, sample_clocks=True, with_e2e=True):
    """Times one BASELINE config on this rank's GPU (weak scaling: every rank owns a full per-GPU batch).  Returns a
    dict with the device-timed step (a CUDA graph WITHOUT instrumentation), the per-stage times of its instrumented
    twin (same kernels, event-record nodes at the stage edges), K1 / K4 alone, the end-to-end number and the clocks."""
    torch, td, world, rank, local_rank, dev, group = (env[k] for k in ("torch", "td", "world", "rank", "local_rank", "dev", "group"))
    from ubpl_b200 import _lib, ops, pipeline, synth
    from ubpl_b200 import dist as ubpl_dist
    c = CONFIGS[cfgname]
    B, K, M, S, J, H, W = (c[k] for k in "BKMSJHW")
    d = synth.make_batch(B=B, K=K, J=J, H=H, W=W, M=M, S=S, seed=1388, rank=rank, device=dev)
    dec = ops.decode_coeffs(d["center"], d["scale"], [H, W])
    w = pipeline.nega_weights(d["islabeled"], 1.0)
    cfg = pipeline.StepConfig(select=c["select"], distThrMax=DIST_THR_MAX,
                              prefetch_student=os.environ.get("UBPL_BENCH_PREFETCH", "1") != "0")
    shapes = json.load(open(os.path.join(ROOT, "ubpl-poseestimation_b200", "hg_param_shapes.json")))[c["hg"]]
    g = torch.Generator(device=dev).manual_seed(5)
    params = [torch.randn(*s, generator=g, device=dev) * 0.02 for s in shapes]
    emas = [torch.randn(*s, generator=g, device=dev) * 0.02 for s in shapes]
    plan = ops.EmaPlan(params, emas)
    n_params = plan.n_elems
    alpha = min(1 - 1 / (3 + 1), 0.999)              # args.epo = 3
    stats = torch.zeros(4, dtype=torch.int64, device=dev)

    p2p_ok = False
    if group is not None and c["select"] == "quantile":
        if os.environ.get("UBPL_BENCH_P2P", "1") != "0":
            # the selector all-reduces its digit histograms over NVLink peer memory inside its one kernel; when the
            # IPC mapping is not possible the NCCL selector (eager K2/K3 stages) stays in use
            if ubpl_dist.p2p_ready(group):
                ubpl_dist.destroy_p2p()
            p2p_ok = ubpl_dist.init_p2p(group, max_items=max(B * J, 1))
        if not p2p_ok:
            ubpl_dist.init_nccl(group)             # the library's own communicator for the histogram all-reduce
    # the EMA rides INSIDE K1's launch (ubpl_warp_decode_k2_ema: the warps that have run out of maps do it while the last
    # maps are decoded -- c2 165.9 vs 167.9 us/step, c3 109.0 vs 114.4 against a separate EMA kernel forked beside K1,
    # which in fact runs in front of it: K1's CTAs do not share an SM); on the multi-GPU quantile path it runs beside
    # the one-CTA selector, whose cross-GPU wait it fills
    overlap = {"0": False, "1": "k1", "k1": "k1", "slow": "k1", "k2": "k2", "k3": "k3", "tail": "tail"}[
        os.environ.get("UBPL_BENCH_OVERLAP_EMA", "k2" if (c["select"] == "quantile" and world > 1) else "tail")]
    gmode = os.environ.get("UBPL_BENCH_GRAPH", "single")
    bufs = dict(teacher=d["teacher"], student=d["student"], theta=d["theta"], flip=d["flip"])
    mk = lambda instrument: pipeline.GraphedStep(bufs["teacher"], bufs["student"], bufs["theta"], bufs["flip"], dec, w, cfg,
                                                 group=group, stats=stats, ema=plan, alpha=alpha, overlap_ema=overlap,
                                                 mode=gmode, instrument=instrument)
    lean_ok = os.environ.get("UBPL_BENCH_LEAN", "1") != "0"
    gstep = mk(not lean_ok)                                   # the step that is timed: no event nodes inside
    single = gstep.mode == "single"
    gtwin = mk(True) if (single and lean_ok) else gstep      # its instrumented twin: the per-stage times
    r = gstep.state

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    stage_events = []

    def step(timed=False):
        evs = {}

        def mark(name):
            if timed and not single:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs[name] = e
        gstep.run(timer=mark)
        if timed and not single:
            stage_events.append(evs)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = None
    if sample_clocks:
        sampler = ClockSampler(local_rank)
        sampler.start()
        sampler.wait_first()
    # launches per step: count them on one eager (un-captured) step
    stats.zero_()
    _lib.reset_launch_count()
    pipeline.pseudo_label_step(d["teacher"], d["student"], d["theta"], d["flip"], dec, w, cfg, group=group)
    plan.step(alpha)
    launches_per_step = _lib.launch_count()
    # the library carries the EMA inside K1's launch up to this many bytes of maps, above it the EMA follows as its own launch
    ema_in_k1 = 4 * H * W * J * M * K * B <= (int(os.environ.get("UBPL_K1_EMA_MAX_MB", "1536")) << 20)
    if overlap == "tail" and cfg.fuse_k12 and M <= 2 and ema_in_k1:
        launches_per_step -= 1                               # the timed step has no EMA launch: K4 rides in K1's
    stats.zero_()
    t_load0 = time.time()
    ms_total = _timed_replays(lambda: step(timed=True), args.steps, barrier, torch)
    slow_frac = float(stats[0]) / max(1.0, float(stats[2]))
    # per-stage device times: event-record nodes inside the twin's graph, read after each of 32 replays of the
    # same loop (a read needs a sync, which must stay out of the timed region)
    stage_samples = []
    if single:
        for _ in range(32):
            gtwin.run()
            torch.cuda.synchronize()
            stage_samples.append(gtwin.stage_ms())
        ms_twin = _timed_replays(lambda: gtwin.run(), min(args.steps, 50), barrier, torch) / min(args.steps, 50)
    else:
        ms_twin = None
    if sampler is not None:
        while time.time() - t_load0 < 0.6:          # keep the load running until nvidia-smi has had >= 0.6 s of it
            for _ in range(20):
                step()
            torch.cuda.synchronize()
        clocks = sampler.stop(t_load0 + 0.05, time.time())
    else:
        clocks = None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = world * B / (ms_step * 1e-3)

    def savg(n):
        return sum(sm[n] for sm in stage_samples) / len(stage_samples)

    def avg(a, b):
        return sum(ev[a].elapsed_time(ev[b]) for ev in stage_events) / len(stage_events)
    if single:
        k1_ms, k2_ms, k3_ms = savg("k1"), savg("k2"), savg("k3")
        k4_inline_ms = savg("k4") if "k4" in stage_samples[0] else None
    else:
        k1_ms, k2_ms, k3_ms = avg("k1_0", "k1_1"), avg("k2_0", "k2_1"), avg("k3_0", "k3_1")
        k4_inline_ms = avg("k4_0", "k4_1") if "k4_0" in stage_events[0] else None

    def alone(fn):                                    # a stage on its own, graph replays back to back
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            fn()
        torch.cuda.synchronize()
        return _timed_replays(gg.replay, 20, lambda: torch.cuda.synchronize(), torch) / 20
    k4_ms = alone(lambda: plan.step(alpha))
    st1 = dict(gstep.state)
    k1_alone_ms = alone(lambda: pipeline.stage_k1(st1, None, cfg))
    # the fused launch (K1 with K4 done by the warps that run out of maps) on its own
    k1_ema_alone_ms = None
    if gstep.overlap_ema == "tail":
        st2 = dict(gstep.state)
        k1_ema_alone_ms = alone(lambda: pipeline.stage_k1(st2, None, cfg, ema=plan, alpha=alpha))
        if not st2.get("ema_done"):
            k1_ema_alone_ms = None

    out = dict(cfg=cfg, c=c, value=value, ms_step=ms_step, ms_twin=ms_twin, k1_ms=k1_ms, k2_ms=k2_ms, k3_ms=k3_ms,
               k4_inline_ms=k4_inline_ms, k4_ms=k4_ms, k1_alone_ms=k1_alone_ms, k1_ema_alone_ms=k1_ema_alone_ms, ema_in_k1=ema_in_k1, launches=launches_per_step * args.steps,
               slow_frac=slow_frac, selected_frac=float(r["enable"].float().mean()), clocks=clocks, n_params=n_params,
               gstep=gstep, overlap=gstep.overlap_ema, single=single, lean=lean_ok and single, p2p_ok=p2p_ok, stats=stats,
               bufs=bufs, data=d, step=step, barrier=barrier)
    gstep.check()                                    # device status words (K2 hand-off, peer-memory selector)

    if with_e2e:
        # ---- end-to-end: host buffers in, host scalars out, copies inside the timed region -----------------------
        pin_local_numa(local_rank, torch)
        host = {k: d[k].cpu().pin_memory() for k in ("teacher", "student", "theta")}
        host["flip"] = gstep.state["flip"].cpu().pin_memory()
        devbuf = {k: gstep.state[k] for k in host}               # the graph reads its inputs from these tensors
        h2d = sum(v.numel() * v.element_size() for v in host.values())

        # Two device input sets, each with its own captured step: the host->device copy of step i+1 (copy stream) runs
        # beside the chain and the device->host read of step i (what pipeline.stream_chunks does for batches that do not
        # fit one GPU).  Every step still moves its 587 MB in and its scalars out, and every result is read on the host --
        # one step behind the launches.  UBPL_BENCH_E2E_PIPELINE=0: copy -> chain -> read, strictly one after the other.
        pipelined = os.environ.get("UBPL_BENCH_E2E_PIPELINE", "1") != "0"
        steps2 = [gstep]
        if pipelined:
            bufs2 = {k: torch.empty_like(devbuf[k]) for k in host}
            steps2.append(pipeline.GraphedStep(bufs2["teacher"], bufs2["student"], bufs2["theta"], bufs2["flip"], dec, w, cfg,
                                               group=group, stats=None, ema=plan, alpha=alpha, overlap_ema=overlap, mode=gmode,
                                               instrument=False))
        nset = len(steps2)
        copy_s, comp = torch.cuda.Stream(), torch.cuda.current_stream()
        filled = [torch.cuda.Event() for _ in range(nset)]
        freed = [torch.cuda.Event() for _ in range(nset)]
        done = [torch.cuda.Event() for _ in range(nset)]
        res = [torch.empty(6, dtype=torch.float64).pin_memory() for _ in range(nset)]
        for ev in freed:
            ev.record(comp)

        def upload(i):
            sidx = i % nset
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(freed[sidx])                   # the step that read this set has finished
                for k in host:
                    steps2[sidx].state[k].copy_(host[k], non_blocking=True)
                filled[sidx].record(copy_s)

        def launch(i):
            sidx = i % nset
            comp.wait_event(filled[sidx])
            rr = steps2[sidx].run()
            res[sidx].copy_(torch.cat([rr["summary"], rr["grad_scale"].double(), rr["count"].double()]), non_blocking=True)
            freed[sidx].record(comp)
            done[sidx].record(comp)

        def e2e_run(n):
            upload(0)
            for i in range(n):
                if i + 1 < n and nset > 1:
                    upload(i + 1)                                # beside the chain of step i
                launch(i)
                if nset == 1:
                    done[0].synchronize()                        # strictly serial: read the result, then the next copy
                    if i + 1 < n:
                        upload(i + 1)
                elif i >= 1:
                    done[(i - 1) % nset].synchronize()           # the host reads step i-1's scalars
            done[(n - 1) % nset].synchronize()
            return res[(n - 1) % nset].clone()
        o = e2e_run(2)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 10))
        o = e2e_run(n_e2e)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        out.update(e2e_val=world * B / (float(t) * 1e-3), e2e_ms=float(t), h2d=h2d, d2h=o.numel() * o.element_size())
    return out


def pin_local_numa(local_rank, torch):
    """Binds this process to the CPUs next to its GPU before the pinned host buffers are allocated (first touch puts
    the pages on that NUMA node), so that N ranks do not all stage their H2D copies through node 0."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:00.0/local_cpulist" % (dom, bus)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
    except Exception:
        pass


def sensitivity_block(m, args, env):
    """K1's cost depends on the data (maps it can prune vs maps that need every output pixel): the same step on
    (a) the synthetic set with EVERY all-negative map pure noise (noise_only_frac = 1.0, the pre-5de460d data) and
    (b) the worst case, every teacher map structure-less noise.  Timed in this run, device resident."""
    torch = env["torch"]
    from ubpl_b200 import synth
    c, d, stats, step, barrier = m["c"], m["data"], m["stats"], m["step"], m["barrier"]
    B, K, M, S, J, H, W = (c[k] for k in "BKMSJHW")
    keep = d["teacher"].clone()
    out = {"noise_only_frac=%.1f (bench default)" % 0.2: {"exhaustive_decode_frac": m["slow_frac"], "ms_per_step": m["ms_step"],
                                                          "k1_standalone_ms": m["k1_alone_ms"]}}
    from ubpl_b200 import pipeline
    variants = [("noise_only_frac=1.0", lambda: synth.make_batch(B=B, K=K, J=J, H=H, W=W, M=M, S=S, seed=1388, rank=env["rank"],
                                                                 device=env["dev"], noise_only_frac=1.0)["teacher"]),
                ("worst case: every map white noise", lambda: torch.randn_like(keep) * 0.02)]
    for name, make in variants:
        d["teacher"].copy_(make())
        for _ in range(3):
            step()
        stats.zero_()
        n = max(5, min(args.steps, 30))
        ms = _timed_replays(step, n, barrier, torch) / n
        frac = float(stats[0]) / max(1.0, float(stats[2]))
        st1 = dict(m["gstep"].state)
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            pipeline.stage_k1(st1, None, m["cfg"])
        torch.cuda.synchronize()
        k1 = _timed_replays(gg.replay, 10, lambda: torch.cuda.synchronize(), torch) / 10
        out[name] = {"exhaustive_decode_frac": frac, "ms_per_step": ms, "k1_standalone_ms": k1,
                     "value": env["world"] * B / (ms * 1e-3)}
    d["teacher"].copy_(keep)
    return out


def ema_placement_block(m, args, env):
    """Where the step's EMA runs.  The timed step carries it INSIDE K1's launch (overlap "tail"): bit-identical
    arithmetic, but in the reference's loop update_ema_variables runs after optimizer.step() (projects/MT_UBPL.py:338),
    i.e. before the NEXT iteration's teacher forward pass -- riding in the next K1 delays it past that forward pass
    (the teacher then lags one more update).  The same step with the EMA as a launch of its own, which is what an
    unchanged driver gets through install(), is timed here beside it."""
    torch = env["torch"]
    from ubpl_b200 import pipeline
    g0, bufs, barrier = m["gstep"], m["bufs"], m["barrier"]
    if g0.ema is None or g0.overlap_ema != "tail":
        return None
    g = pipeline.GraphedStep(bufs["teacher"], bufs["student"], bufs["theta"], bufs["flip"], g0.state["dec"], g0.state["sample_w"],
                             m["cfg"], group=g0.group, ema=g0.ema, alpha=g0.alpha, overlap_ema=False, mode=g0.mode,
                             instrument=False)
    for _ in range(3):
        g.run()
    n = max(5, min(args.steps, 50))
    ms = _timed_replays(g.run, n, barrier, torch) / n
    B = m["c"]["B"]
    return {"inside_k1_launch": {"ms_per_step": m["ms_step"], "value": m["value"]},
            "own_launch_after_k3": {"ms_per_step": ms, "value": env["world"] * B / (ms * 1e-3)},
            "note": "`value` is the first; the second keeps the reference's order (update_ema_variables as its own call)"}


def ops_block(cfgname, env, peak):
    """The kernels behind the criteria the reference's drivers call every step (utils/losses.py:169-210 via
    ubpl_dense_mse, utils/process.py:19-31 via ubpl_features_cov) and the materialised back-warp
    (utils/augment.py:37-47): GB/s of algorithmic bytes and fraction of the measured HBM peak, CUDA-graph replays."""
    torch = env["torch"]
    from ubpl_b200 import ops
    c = CONFIGS[cfgname]
    B, M, S, J, H, W = c["B"], max(c["M"], 1), c["S"], c["J"], c["H"], c["W"]
    dev = env["dev"]
    g = torch.Generator(device=dev).manual_seed(9)
    res = {}

    def bench(fn, nbytes, n=20):
        for _ in range(3):
            fn()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            fn()
        torch.cuda.synchronize()
        ms = _timed_replays(gg.replay, n, lambda: torch.cuda.synchronize(), torch) / n
        return {"ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "algorithmic_bytes": nbytes}
    pred = torch.rand(B, S, J, H, W, generator=g, device=dev)
    for Mt in sorted({1, 2, M}):
        tgt = torch.rand(Mt, B, S, J, H, W, generator=g, device=dev)
        # JointPseudoLoss3: student maps read once, the last-stack teacher maps of the Mt teachers read once, gradient
        # written once: 4*HW*J*B*(S + Mt + S)
        res["dense_mse_kernel as JointPseudoLoss3 (M=%d, %s sizes)" % (Mt, cfgname)] = bench(
            lambda: ops.dense_mse(pred, tgt[:, :, -1], mask_mode=1, thr=0.95), 4 * H * W * J * B * (S + Mt + S))
        del tgt
    f1 = torch.randn(B // 2, S, 256, 32, 32, generator=g, device=dev)
    f2 = torch.randn(B // 2, S, 256, 32, 32, generator=g, device=dev)
    res["features_cov_kernel fwd+bwd (%d labeled rows, 16 B/element)" % (B // 2)] = bench(
        lambda: ops.features_cov(f1, f2), 16 * f1.numel())
    del f1, f2
    hm = torch.rand(B, J, H, W, generator=g, device=dev)
    th = torch.zeros(B, 2, 3, device=dev)
    th[:, 0, 0] = 0.8
    th[:, 1, 1] = 0.8
    fl = torch.zeros(B, dtype=torch.uint8, device=dev)
    res["warp_materialize_kernel (affine_back2, one view, 8 B/texel)"] = bench(
        lambda: ops.warp_materialize(hm, th, fl), 8 * hm.numel())
    return res


def run_ours(args):
    import torch
    import ubpl_b200  # noqa: F401

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group, td = None, None
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)
        group = td.group.WORLD
    env = dict(torch=torch, td=td, world=world, rank=rank, local_rank=local_rank, dev=dev, group=group)
    c = CONFIGS[args.config]
    B, K, M, S, J, H, W = (c[k] for k in "BKMSJHW")
    m = measure_config(args.config, args, env)
    cfg = m["cfg"]
    k1_ms, k2_ms, k3_ms, k4_ms = m["k1_ms"], m["k2_ms"], m["k3_ms"], m["k4_ms"]
    bytes_sample = algorithmic_bytes_per_sample(c)
    k1_bytes = 4 * H * W * J * M * K * B
    k3_bytes = 4 * H * W * J * (2 * S + 1) * B
    ema_bytes = 12 * m["n_params"]
    peak, peak_src = measured_peaks()
    chain_gbs = bytes_sample * B / ((k1_ms + k2_ms + k3_ms) * 1e-3) / 1e9
    ov = m["overlap"]
    # with the EMA inside K1's launch (overlap "tail") the launch's algorithmic bytes are K1's maps plus the EMA's 12 B per
    # parameter; the K1-only figure (the EMA's bytes not counted, its time counted) is kept beside it
    fused = ov == "tail" and m.get("k1_ema_alone_ms") is not None
    launch_bytes = k1_bytes + (ema_bytes if fused else 0)
    roof = {"bound": "hbm", "kernel": "warp_decode_kernel (K1: %d maps of %d B per launch%s)" %
                                      (M * K * B * J, 4 * H * W,
                                       (" + K4: the EMA of %d parameters in its tail" if m["ema_in_k1"] else
                                        " + K4: the EMA of %d parameters as a launch of its own right behind it (the launch is too long for the fused instance to pay)")
                                       % m["n_params"] if fused else ""),
            "achieved": launch_bytes / (k1_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": launch_bytes / (k1_ms * 1e-3) / 1e9 / peak,
            "achieved_k1_bytes_only": k1_bytes / (k1_ms * 1e-3) / 1e9, "frac_k1_bytes_only": k1_bytes / (k1_ms * 1e-3) / 1e9 / peak,
            "launch_bytes": launch_bytes,
            "traffic": NCU_TRAFFIC.get(args.config), "traffic_source": "from profile (ncu --set full capture under profiles/, not measured in this run)",
            "peak_source": peak_src,
            "stages_ms": {("k1_warp_decode_with_k4_ema_overlapped" if ov == "k1" else
                           "k1_warp_decode_with_k4_ema_in_its_tail" if ov == "tail" else "k1_warp_decode"): k1_ms,
                          ("k2_uncertainty_select_with_k4_ema_overlapped" if ov == "k2" else "k2_uncertainty_select"): k2_ms,
                          ("k3_render_mse_with_k4_ema_overlapped" if ov == "k3" else "k3_render_mse"): k3_ms,
                          "k4_ema_in_step": m["k4_inline_ms"], "k4_ema_standalone": k4_ms, "k1_standalone": m["k1_alone_ms"]},
            "k1_standalone_frac": k1_bytes / (m["k1_alone_ms"] * 1e-3) / 1e9 / peak,
            "k1_with_ema_standalone_ms": m.get("k1_ema_alone_ms"),
            "k1_with_ema_standalone_frac": (launch_bytes / (m["k1_ema_alone_ms"] * 1e-3) / 1e9 / peak) if fused else None,
            "stages_note": ("`value` times the step as ONE CUDA graph without instrumentation (%.4f ms/step); the stage times are "
                            "the event-record nodes inside its instrumented twin (same kernels; each event node adds ~1.3 us: "
                            "%.4f ms/step), mean of 32 replays right after the timed region"
                            % (m["ms_step"], m["ms_twin"]) if (m["single"] and m["lean"]) else
                            "stage edges are event-record nodes inside the step's single CUDA graph" if m["single"] else
                            "stage edges are CUDA events recorded around each stage graph in every timed step"),
            "k2_fused_into_k1": bool(cfg.fuse_k12 and M <= 2),
            "stages_gbs": {"k1": k1_bytes / (k1_ms * 1e-3) / 1e9, "k3": k3_bytes / (k3_ms * 1e-3) / 1e9,
                           "k4": ema_bytes / (k4_ms * 1e-3) / 1e9, "chain_k1_k3": chain_gbs},
            "chain_frac_of_peak": chain_gbs / peak, "chain_frac_of_8TBs": chain_gbs / 8000.0,
            "step_gbs_with_ema": (bytes_sample * B + ema_bytes) / (m["ms_step"] * 1e-3) / 1e9,
            "step_frac_of_peak_with_ema": (bytes_sample * B + ema_bytes) / (m["ms_step"] * 1e-3) / 1e9 / peak,
            "step_frac_of_8TBs_with_ema": (bytes_sample * B + ema_bytes) / (m["ms_step"] * 1e-3) / 1e9 / 8000.0,
            "algorithmic_bytes_per_sample": bytes_sample}

    line = {
        "metric": METRIC, "value": m["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": m["ms_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.config), **{
            "ema_params": m["n_params"], "l2": "inputs (%.0f MB/step) larger than L2" % (bytes_sample * B / 1e6),
            "selected_frac": m["selected_frac"], "exhaustive_decode_frac": m["slow_frac"], "noise_only_frac": 0.2,
            "k1_copies_in_flight_cap": int(os.environ.get("UBPL_K1_INFLIGHT", "8")),
            "launch": ("1 CUDA graph per step" + (" (no event nodes in the timed graph)" if m["lean"] else "")
                       if m["single"] else "stage graphs (%s)" % ", ".join(
                           n + (":eager" if n in m["gstep"].eager else ":graph") for n in m["gstep"].order))
                      + (("; EMA inside K1's launch (done by the warps that have run out of maps)" if m["ema_in_k1"] else
                        "; EMA as its own launch right behind K1 (ubpl_warp_decode_k2_ema decides by the launch's size)") if ov == "tail" else
                         "; EMA forked onto a side stream beside %s" % {"k1": "K1", "k2": "the selector", "k3": "K3"}[ov]
                         if ov else "; EMA after K3"),
            "selector": ("fixed rule in K1's epilogue" if c["select"] == "fixed" and cfg.fuse_k12 and M <= 2 else
                         "one-kernel quantile selector" + (", histograms all-reduced over NVLink peer memory" if m["p2p_ok"] else "")
                         if (world == 1 or m["p2p_ok"]) and c["select"] == "quantile" else
                         "NCCL histogram all-reduce" if c["select"] == "quantile" else "k2 kernels")}),
        "roofline": roof, "cpu_baseline": None,
        "e2e": {"value": m["e2e_val"], "unit": "samples/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "ms_per_step": m["e2e_ms"],
                "note": "PCIe-bound: the step's inputs (%.0f MB) cross the host link every step (55.4 GB/s on this box's link = 10.6 ms, "
                        "tools/h2d_micro.py); the copy of step i+1 runs beside the chain and the result read of step i "
                        "(two device input sets; UBPL_BENCH_E2E_PIPELINE=0 serialises them: same number)" % (m["h2d"] / 1e6)},
        "gpu_launches": m["launches"], "clocks": m["clocks"],
    }
    if world == 1 and not args.no_extras:
        line["ema_placement"] = ema_placement_block(m, args, env)
        line["sensitivity"] = sensitivity_block(m, args, env)
        line["ops"] = ops_block(args.config, env, peak)
    del m
    torch.cuda.empty_cache()
    if args.config != "c4" and not args.no_extras:
        # the one config with a collective (global-quantile threshold): timed in the same run at EVERY N (at N = 1 the
        # selector has no peer to talk to) so that the driver's scaling record covers the selector's cross-GPU step and
        # its own N = 1 denominator; `value` above stays the headline config
        import copy
        a2 = copy.copy(args)
        a2.steps, a2.warmup = max(10, min(args.steps, 100)), max(3, min(args.warmup, 10))
        mc = measure_config("c4", a2, env, sample_clocks=False, with_e2e=False)
        cc = CONFIGS["c4"]
        line["collective"] = {
            "workload": "c4: " + cc["desc"], "value": mc["value"], "unit": "samples/s", "ms_per_step": mc["ms_step"],
            "steps": a2.steps, "per_gpu_batch": cc["B"], "k1_stage_us": mc["k1_ms"] * 1e3, "k2_stage_us": mc["k2_ms"] * 1e3,
            "k3_stage_us": mc["k3_ms"] * 1e3,
            "selector": ("digit histograms all-reduced over NVLink peer memory inside one kernel (ubpl_select_quantile_fused)"
                         if mc["p2p_ok"] else "one-kernel selector, single GPU (no exchange)" if world == 1 else
                         "NCCL histogram all-reduce (ubpl_select_quantile_dist)"),
            "n1_reference": C4_N1,
            "vs_n1": (mc["value"] / C4_N1["value"]) if C4_N1.get("value") else None,
            "launch": "1 CUDA graph per step" if mc["single"] else "stage graphs + eager NCCL stage"}
        del mc
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            kind = cpu_kind()
            procs = max(1, min(os.cpu_count() or 1, 64))
            chain = CpuChain(args.config, procs, kind)
            per_proc = 1 if H * W * J * K * M > 64 * 64 * 14 * 8 * 4 else 4
            n = dt = 0
            t_cpu0 = time.perf_counter()
            while dt < 10.0 and time.perf_counter() - t_cpu0 < 60.0:          # ~10 s of CPU work
                nn, d_t = chain.step(per_proc)
                n += nn
                dt += d_t
            chain.close()
            ema_s = cpu_ema_seconds(c, kind)
            line["cpu_baseline"] = {"value": n / (dt + ema_s * n / B), "unit": "samples/s", "cores": procs, "kind": kind,
                                    "sample": cpu_sample_note(kind, procs * per_proc, procs, per_proc, args.config)
                                    + "; %d samples in %.1f s" % (n, dt)}
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sensitivity / ops / collective blocks")
    ap.add_argument("--ref-samples-per-proc", type=int, default=0, help="CPU arm: samples per process per step (0 = by config)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
