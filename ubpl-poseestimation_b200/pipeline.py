"""The fused pseudo-label chain, device resident, no host sync:

    K1  back-warp + flip + arg-max decode of every teacher view      (augment.py:37-47, process.py:330-336)
    K2  per-joint dispersion -> selection mask                        (evaluation.py:40-58 / business.py:109-217)
    K3  Gaussian target render + masked joint-MSE forward + gradient  (process.py:253-278, losses.py:8-29)

`pseudo_label_step` composes the reference's library functions in the order of SURVEY 3.4:
kps_fromHeatmap_mul on the back-warped maps of every teacher -> (M=1) uncertainty_fromDistance
dispersion, or (M=2) assess_pseudo_unc2 on [predsMean_1, predsMean_2, preds_mean] ->
filter_pseudo2 (global quantile) or the fixed rule of pseudo_filter_mixUnc -> Dataset.update
(coords + enable) -> kps_heatmap -> JointMSELoss(useKPsGate, useSampleWeight) with the
`getSampleWeight_nega` weights (unlabeled rows get pseudoWeight) -> weight * sum / n
(MT_UBPL.py:266) and its gradient."""
import os
from dataclasses import dataclass

import torch

from . import ops


@dataclass
class StepConfig:
    select: str = "fixed"            # "fixed" (business.py:237-261) or "quantile" (business.py:173-217)
    distThrMax: float = 1.0          # fixed rule: unc <= 1-exp(-3*distThrMax/5)
    reliableThr: float = 0.0
    reliablePCT: float = 0.5
    reliableDistMin: float = 1.0
    pseudoWeight: float = 1.0        # projects/tools.py:24-31
    lossWeight: float = 1.0          # args.poseWeight, MT_UBPL.py:266
    stride: float = 4.0              # inpRes / outRes, process.py:255
    sigma: float = 3.0               # kernelSize * sigma, process.py:258
    want_target: bool = True         # materialise the rendered targets (counted in the byte model)
    want_grad: bool = True
    fuse_k2: bool = True             # one-launch K2 for the M=1 fixed-threshold path
    fuse_k12: bool = True            # the per-joint part of K2 runs in K1's epilogue (no K2 launch on the fixed path)
    fuse_sum: bool = True            # the loss reduction runs in K3's last CTA (no loss_finalize launch)
    swap_perm: object = None         # optional [J] left/right exchange for flipped views (flip_back,
                                     # utils/udaap/transforms.py:20-57); None = the reference's live path
    prefetch_student: bool = True    # K1's idle warps pull the student maps (K3's input) into L2 during K1's tail


def nega_weights(islabeled, pseudoWeight):
    """projects/tools.py:24-31 getSampleWeight_nega: labeled rows 0, unlabeled rows pseudoWeight."""
    return torch.where(islabeled.bool(), torch.zeros((), device=islabeled.device),
                       torch.full((), float(pseudoWeight), device=islabeled.device)).to(torch.float32)


def stage_k1(st, stats=None, cfg=None, ema=None, alpha=None, alpha_from_device=True):
    """K1: back-warp + flip + arg-max decode of every (model, view) map, each read from HBM once.  With one
    teacher (and cfg.fuse_k12) the per-joint dispersion -- and on the fixed path the whole selection -- is
    computed by the warp that decodes the last view of a joint, inside the same launch.  ema (an ops.EmaPlan; alpha
    read from its device buffer unless alpha_from_device=False): K4 inside the same launch, done by the warps that
    have run out of maps (fused path only; st["ema_done"] = True when the EMA went along)."""
    teacher, theta, flip, dec = st["teacher"], st["theta"], st["flip"], st["dec"]
    M, K, B, J, H, W = teacher.shape
    st.pop("k12", None)
    flat = teacher.stride(0) == K * teacher.stride(1)              # [M,K,...] is one run of M*K maps
    if (cfg is not None and cfg.fuse_k12 and cfg.select in ("fixed", "quantile") and
            ((M == 1 and K <= 32) or (M == 2 and K <= 16 and flat))):
        S = st["student"].shape[1]
        sH, sW = st["student"].shape[-2:]
        mode = (2 if cfg.select == "fixed" else 1) + (2 if M == 2 else 0)
        if M == 1:
            maps, th, fl = teacher[0], theta, flip
        else:                                                      # both teachers see the same K views
            maps = (teacher.view(M * K, B, J, H, W) if teacher.is_contiguous() else
                    teacher.as_strided((M * K, B, J, H, W), (teacher.stride(1),) + tuple(teacher.stride()[2:])))
            th = theta.unsqueeze(0).expand(M, K, B, 2, 3).reshape(M * K, B, 2, 3)
            fl = flip.unsqueeze(0).expand(M, K, B).reshape(M * K, B)
        r = ops.warp_decode_k2(maps, th, fl, dec, mode, S=S, img_h=int(sH * cfg.stride), img_w=int(sW * cfg.stride),
                               stride=cfg.stride, sigma=cfg.sigma, distThrMax=cfg.distThrMax, stats=stats,
                               swap_perm=cfg.swap_perm, prefetch=st["student"] if cfg.prefetch_student else None,
                               ema=ema, alpha=alpha, alpha_from_device=alpha_from_device and ema is not None)
        st["xy"], st["max"], st["idx"] = r["xy"].view(M, K, B, J, 2), r["max"].view(M, K, B, J), r["idx"].view(M, K, B, J)
        st["k12"] = r
        st["ema_done"] = ema is not None
        return st
    st["ema_done"] = False
    perm = cfg.swap_perm if cfg is not None else None
    if flat:
        dec_out = ops.warp_decode(teacher.view(M * K, B, J, H, W) if teacher.is_contiguous() else
                                  teacher.as_strided((M * K, B, J, H, W), (teacher.stride(1),) + tuple(teacher.stride()[2:])),
                                  theta.unsqueeze(0).expand(M, K, B, 2, 3).reshape(M * K, B, 2, 3),
                                  flip.unsqueeze(0).expand(M, K, B).reshape(M * K, B), dec, stats=stats, want_idx=True,
                                  swap_perm=perm)
        st["xy"] = dec_out["xy"].view(M, K, B, J, 2)
        st["max"] = dec_out["max"].view(M, K, B, J)
        st["idx"] = dec_out["idx"].view(M, K, B, J)
    else:
        outs = [ops.warp_decode(teacher[m], theta, flip, dec, stats=stats, swap_perm=perm) for m in range(M)]
        st["xy"] = torch.stack([o["xy"] for o in outs])
        st["max"] = torch.stack([o["max"] for o in outs])
        st["idx"] = torch.stack([o["idx"] for o in outs])
    return st


def stage_k2(st, cfg, group=None):
    """K2: per-joint dispersion -> selection mask -> visibility gate and open-gate count."""
    xy = st["xy"]
    M, K, B, J, _ = xy.shape
    S = st["student"].shape[1]
    H, W = st["student"].shape[-2:]
    stride = cfg.stride
    img_h, img_w = int(H * stride), int(W * stride)
    k12 = st.get("k12")
    if k12 is not None and cfg.select == "fixed":            # everything was computed in K1's epilogue
        st.update(kps=k12["mean"], dist=k12["dist"], legal=k12["legal"], gate=k12["gate"], grad_scale=None,
                  count=k12["count"], enable=k12["enable"], counts=k12["counts"])
        return st
    if k12 is not None:                                      # quantile: the selector is the only K2 launch left
        kps, dist, legal = k12["mean"], k12["dist"], k12["legal"]
        _select_and_gate(st, cfg, group, kps, dist, legal, J, S, img_h, img_w)
        return st
    if cfg.fuse_k2 and M == 1 and cfg.select == "fixed":
        k2 = ops.k2_view_fixed(xy[0], cfg.distThrMax, S, img_h, img_w, stride, cfg.sigma)
        st.update(kps=k2["mean"], dist=k2["dist"], legal=k2["legal"], gate=k2["gate"], grad_scale=None,
                  count=k2["count"], enable=k2["enable"], counts=k2["counts"])
        return st
    if M == 1:
        vd = ops.view_dispersion(xy[0], sentinel_illegal=True)
        kps, dist, legal = vd["mean"], vd["dist"], vd["legal"]
        st.update(unc32=vd["unc32"], max_bits=vd["max_bits"])
    elif M == 2:
        vd1, vd2 = ops.view_dispersion(xy[0]), ops.view_dispersion(xy[1])
        ad = ops.assess_dual(vd1["mean"], vd2["mean"], None, xy[0], xy[1])
        kps, dist, legal = ad["coord32"], ad["extDist"], ad["legal"]
        st.update(assess=ad)
    else:
        raise ValueError("pseudo_label_step supports M = 1 (mean teacher) or M = 2 (dual teachers)")
    _select_and_gate(st, cfg, group, kps, dist, legal, J, S, img_h, img_w)
    return st


def _select_and_gate(st, cfg, group, kps, dist, legal, J, S, img_h, img_w):
    """selection mask -> visibility gate, open-gate count and gradient scale (one launch on the quantile path
    when the fused selector is in use, else selector + gate_prepare)."""
    if cfg.select == "fixed":
        sel = ops.select_fixed(dist, legal, J, cfg.distThrMax)
    elif cfg.select == "quantile":
        sel = ops.select_quantile(dist, legal, J, cfg.reliableThr, cfg.reliablePCT, cfg.reliableDistMin, group=group,
                                  gate=(kps, S, img_h, img_w, cfg.stride, cfg.sigma, cfg.lossWeight))
    else:
        raise ValueError("select must be 'fixed' or 'quantile'")
    if sel.get("gate_fused"):
        gate, grad_scale, count = sel["gate"], sel["grad_scale"], sel["count"]
    else:
        gate, grad_scale, count = ops.gate_prepare(kps, sel["gate"], S, img_h, img_w, cfg.stride, cfg.sigma, cfg.lossWeight)
    st.update(kps=kps, dist=dist, legal=legal, gate=gate, grad_scale=grad_scale, count=count, enable=sel["enable"],
              counts=sel["counts"], sel=sel)


def stage_k3(st, cfg):
    """K3: Gaussian target render + masked joint-MSE forward + gradient in one pass, then the loss reduction."""
    student = st["student"]
    B, S, J, H, W = student.shape
    stride = cfg.stride
    img_h, img_w = int(H * stride), int(W * stride)
    gs = st["grad_scale"]
    r = ops.render_mse(st["kps"], st["gate"], st["sample_w"], student, img_h, img_w, stride, cfg.sigma, grad_scale=gs,
                       want_grad=cfg.want_grad, want_target=cfg.want_target,
                       count_in=st["count"] if gs is None else None, loss_weight=cfg.lossWeight,
                       want_summary=cfg.fuse_sum)
    if gs is None:
        st["grad_scale"] = r["grad_scale"]
    st["gate"] = st["gate"].view(B, J)
    st["enable"] = st["enable"].view(B, J)
    st["summary"] = r["summary"] if cfg.fuse_sum else ops.loss_finalize(r["per_loss"], None, st["gate"])
    st.update(grad=r["grad"], target=r["target"], per_loss=r["per_loss"])
    return st


def pseudo_label_step(teacher, student, theta, flip, dec, sample_w, cfg: StepConfig, group=None, stats=None,
                      timer=None, ema=None, alpha=None):
    """teacher [M,K,B,J,H,W] (last-stack teacher maps of the K augmented views), student
    [B,S,J,H,W], theta [K,B,2,3], flip [K,B], dec [B,4] (ops.decode_coeffs), sample_w [B].
    Returns a dict of DEVICE tensors: summary float64[4] = (loss_sum, #loss>0, #mask>0, #gate>0),
    grad_scale, count, grad, target, gate, kps, enable, dist, decode outputs.  The scalar loss of
    MT_UBPL.py:266 is summary[0] * grad_scale.  No host synchronisation anywhere.
    ema (an ops.EmaPlan) + alpha: the step's mean-teacher update (utils/parameters.py:4-8) goes along -- inside K1's
    launch on the fused path (the warps that run out of maps do it), as a launch of its own behind K1 otherwise.
    Mind the order: the reference updates the teacher after optimizer.step(), before the next forward pass; an update
    that rides in this call lands after the forward passes that produced `teacher` / `student`."""
    mark = timer if timer is not None else (lambda name: None)
    st = dict(teacher=teacher, student=student, theta=theta, flip=flip, dec=dec, sample_w=sample_w)
    mark("k1_0")
    if ema is not None:
        stage_k1(st, stats, cfg, ema=ema, alpha=alpha, alpha_from_device=False)
        if not st.get("ema_done"):
            ema.step(alpha)
    else:
        stage_k1(st, stats, cfg)
    mark("k1_1")
    stage_k2(st, cfg, group)
    mark("k3_0")
    stage_k3(st, cfg)
    mark("k3_1")
    return st


def view_targets(kps, gate, mats, flips, H, W, img_h, img_w, stride=None, sigma=3.0):
    """in-frame targets of the selected pseudo-labels for every augmented student view, on the device.
    kps [B,J,2] canonical image coordinates (st["kps"]), gate [B,J] (st["gate"], becomes the key-point weight),
    mats [V,B,3,3] float64 = AugmentUtils.view_matrix of each (view, sample), flips [V,B].  What Dataset.update +
    the Dataset's own augmentation of the key points + kps_heatmap do per sample and joint on the host
    (datasets/dataset_mds.py:14-25,98-117, utils/process.py:253-278): returns (heatmaps [V,B,J,H,W],
    kps_view [V,B,J,3] with weight *= visibility in the view)."""
    B, J, _ = kps.shape
    V = mats.shape[0]
    k3 = torch.cat([kps.to(torch.float32), gate.reshape(B, J, 1).to(torch.float32)], -1)
    kv = ops.view_kps(k3, mats, flips, img_w)
    hm, kout = ops.render_targets(kv.reshape(-1, 3), H, W, img_h, img_w, stride, sigma)
    return hm.view(V, B, J, H, W), kout.view(V, B, J, 3)


class GraphedStep:
    """The same chain captured into CUDA graphs for fixed shapes and fixed input buffers: one step is only
    ~0.2 ms of GPU time, so eager launches (plus allocator and Python work) would leave the GPU waiting for
    the host.  Inputs are read from the tensors given here (copy new batches into them); outputs live in
    `self.state`.  `ema` is an optional ops.EmaPlan whose update is K4.

    mode "single" (default): ONE graph per step; the stage edges are external CUDA events recorded by
    event-record nodes inside the graph, so `stage_ms()` gives the device time of each stage of the most
    recent replay without splitting the step into several launches.
    mode "stages": one graph per stage (K1, K2, K3, K4), stage edges recorded by the caller's `timer`.
    A stage whose collective cannot be captured (the NCCL selector) runs eagerly and forces "stages"."""

    def __init__(self, teacher, student, theta, flip, dec, sample_w, cfg: StepConfig, group=None, stats=None,
                 ema=None, alpha=None, warmup=3, overlap_ema=True, mode="single", instrument=True):
        self.cfg, self.group, self.ema, self.alpha = cfg, group, ema, alpha
        self.state = dict(teacher=teacher, student=student, theta=theta, flip=flip.to(torch.uint8), dec=dec,
                          sample_w=sample_w)
        self.graphs = {}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # first-call initialisation must not be captured
                self._eager(stats)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pool = None
        self.eager = {}
        # overlap_ema: True / "tail" = K4 INSIDE K1's launch (ubpl_warp_decode_k2_ema): the warps that have run out of maps do
        # the EMA while the last maps are decoded (the fused K1+K2 path only; otherwise K4 follows K1 on the same
        # stream); "k1" = K4 forked onto a side stream beside K1 (it then runs in front of K1: K1's CTAs do not
        # share an SM with other kernels),
        # "k2" = beside the quantile selector (a one-CTA kernel), "k3" = beside K3, False = after K3
        if overlap_ema == "slow":                         # round-1 name: K1 no longer has a second launch
            overlap_ema = "k1"
        self.overlap_ema = (overlap_ema if overlap_ema in ("k1", "k2", "k3") else "tail") if (overlap_ema and ema is not None) else False
        if self.overlap_ema == "k2" and cfg.select != "quantile":
            self.overlap_ema = "tail"                     # the fixed path has no K2 launch to hide behind
        self._side = torch.cuda.Stream() if (self.overlap_ema and self.overlap_ema != "tail") else None
        self.events = {}
        if ema is not None:
            ema.set_alpha(alpha)                          # the captured EMA launch reads alpha from device memory

        self._join_late = mode == "single" and os.environ.get("UBPL_EMA_JOIN", "late") == "late"
        self._pending_join = False

        def forked(fn):
            # K4 is independent of the chain: fork it onto a side stream inside the same graph so that it runs
            # concurrently with the stage (K1 is not bandwidth-saturated on its own).  In the single-graph step the side
            # stream joins at the END of the step (nothing in the chain reads the teacher weights), so the next stage
            # follows its predecessor kernel directly -- which is what lets K3 start with programmatic dependent launch
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self._ema()
            fn()
            if self._join_late:
                self._pending_join = True
            else:
                torch.cuda.current_stream().wait_stream(self._side)

        def join():
            if self._pending_join:
                torch.cuda.current_stream().wait_stream(self._side)
                self._pending_join = False

        def k1_fn():
            if self.overlap_ema == "k1":
                forked(lambda: stage_k1(self.state, stats, cfg))
            elif self.overlap_ema == "tail":
                stage_k1(self.state, stats, cfg, ema=self.ema, alpha=self.alpha)
                if not self.state.get("ema_done"):
                    self._ema()
            else:
                stage_k1(self.state, stats, cfg)

        def k3_fn():
            if self.overlap_ema == "k3":
                forked(lambda: stage_k3(self.state, cfg))
            else:
                stage_k3(self.state, cfg)

        def k2_fn():
            if self.overlap_ema == "k2":
                forked(lambda: stage_k2(self.state, cfg, group))
            else:
                stage_k2(self.state, cfg, group)

        stages = [("k1", k1_fn), ("k2", k2_fn), ("k3", k3_fn)]
        if ema is not None and not self.overlap_ema:
            stages.append(("k4", self._ema))
        from . import dist as _dist
        nccl_eager = (group is not None and cfg.select == "quantile" and not _dist.p2p_ready(group))
        if nccl_eager:
            mode = "stages"
        self.mode = mode
        if mode == "single":
            # an event-record node costs ~1-2 us of device time, so only the edges between stages that launch
            # something are instrumented: on the fixed path with K2 fused into K1 the K2 stage is empty and is
            # folded into the K1 segment
            Mt, Kt = teacher.shape[0], teacher.shape[1]
            k2_empty = (cfg.fuse_k12 and cfg.select == "fixed" and
                        ((Mt == 1 and Kt <= 32) or (Mt == 2 and Kt <= 16 and teacher.stride(0) == Kt * teacher.stride(1))))
            segs = []
            for name, fn in stages:
                if name == "k2" and k2_empty:
                    segs[-1][1].append(fn)
                else:
                    segs.append((name, [fn]))
            # instrument=False: the same graph without the event-record nodes (each costs ~1.3 us of device time and
            # sits between two stages); stage_ms() is then unavailable
            try:
                ev = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(len(segs) + 1)] if instrument else "lean"
            except TypeError:                             # torch without external events
                ev = None
            if ev == "lean":
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for name, fns in segs:
                        for fn in fns:
                            fn()
                    join()
                self.graphs["step"] = g
                self.order = ["step"]
                self.stage_names = [n for n, _ in stages]
                return
            if ev is not None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ev[0].record()
                    for i, (name, fns) in enumerate(segs):
                        for fn in fns:
                            fn()
                        if i == len(segs) - 1:
                            join()                        # the EMA's tail (if any) is charged to the last stage
                        ev[i + 1].record()
                self.graphs["step"] = g
                self.events = {name: (ev[i], ev[i + 1]) for i, (name, _) in enumerate(segs)}
                self.order = ["step"]
                self.stage_names = [n for n, _ in stages]
                return
            self.mode = mode = "stages"
        for name, fn in stages:
            if name == "k3" and "k2" in self.eager:
                self.eager["k3"] = fn             # K2's outputs are re-allocated every step when it runs eagerly
                continue
            if name == "k2" and nccl_eager:
                # the stage with the NCCL all-reduces is launched eagerly (one fused C call enqueues its kernels
                # and collectives); its outputs are re-allocated every step, so K3 stays eager as well
                self.eager["k2"] = fn
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                fn()
            pool = g.pool()
            self.graphs[name] = g
        self.order = [n for n in ("k1", "k2", "k3", "k4") if n in self.graphs or n in self.eager]
        self.stage_names = list(self.order)

    def _ema(self):
        self.ema.step(self.alpha, from_device=True)

    def set_alpha(self, alpha):
        """The EMA factor of the following replays (utils/parameters.py:6: alpha = min(1 - 1/(epo+1), ema_decay)
        changes every epoch).  The captured launch reads {alpha, 1-alpha} from a device buffer, so no re-capture."""
        self.alpha = alpha
        if self.ema is not None:
            self.ema.set_alpha(alpha)

    def check(self):
        """Raises if the K2 epilogue of the most recent replay reported a lost hand-off word (one D2H sync)."""
        k12 = self.state.get("k12")
        if k12 is not None:
            ops.check_status(k12.get("status"))
        if self.group is not None and self.cfg.select == "quantile":
            from . import dist as _dist
            if _dist.p2p_ready(self.group):
                _dist.check_p2p()

    def _eager(self, stats):
        stage_k1(self.state, stats, self.cfg)
        stage_k2(self.state, self.cfg, self.group)
        stage_k3(self.state, self.cfg)
        if self.ema is not None:
            self._ema()

    def run(self, timer=None):
        mark = timer if timer is not None else (lambda name: None)
        for n in self.order:
            mark(n + "_0")
            if n in self.graphs:
                self.graphs[n].replay()
            else:
                self.eager[n]()
            mark(n + "_1")
        return self.state

    def stage_ms(self):
        """mode "single": device milliseconds of every stage of the most recent replay (the caller must have
        synchronised), from the event-record nodes inside the graph."""
        if not self.events:
            return None
        return {n: (self.events[n][0].elapsed_time(self.events[n][1]) if n in self.events else 0.0) for n in self.stage_names}


def stream_chunks(host, dec, sample_w, cfg: StepConfig, chunk, group=None, ema=None, alpha=None, on_chunk=None):
    """A batch that does not fit one GPU (BASELINE config 5: B = 4096, K = 16, J = 32, 128x128 is 180 GB of maps) streamed
    through the chain in chunks of `chunk` samples: two device buffer sets, each with its own captured step, so that the
    host->device copy of chunk i+1 (copy stream, pinned host memory) overlaps the chain of chunk i (compute stream).
    host: dict of PINNED CPU tensors teacher [M,K,B,J,H,W], student [B,S,J,H,W], theta [K,B,2,3], flip [K,B] (uint8);
    dec [B,4] float64 and sample_w [B] may live on the device already (tiny).  The global-quantile selection is per
    chunk (a chunk is the selector's population).  Returns a list with, per chunk, the float64[4] loss summary,
    grad_scale and the open-gate count (device tensors, cloned); on_chunk(i, state) -- if given -- is called on the
    compute stream after chunk i's step, before its buffers are reused (copy the gradient out there).  The EMA
    (`ema`, an ops.EmaPlan) runs once, with the last chunk."""
    B = host["student"].shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    starts = list(range(0, B, chunk))
    if B % chunk:
        raise ValueError("stream_chunks: the batch (%d) must be a multiple of the chunk (%d)" % (B, chunk))
    copy_s = torch.cuda.Stream()
    comp = torch.cuda.current_stream()
    sets, steps, filled, freed = [], [], [], []
    dec = dec.to(dev)
    sample_w = sample_w.to(dev)
    for s in range(2):
        b0 = starts[min(s, len(starts) - 1)]
        bufs = dict(teacher=host["teacher"][:, :, b0:b0 + chunk].to(dev), student=host["student"][b0:b0 + chunk].to(dev),
                    theta=host["theta"][:, b0:b0 + chunk].to(dev), flip=host["flip"][:, b0:b0 + chunk].to(dev).to(torch.uint8),
                    dec=dec[b0:b0 + chunk].clone(), w=sample_w[b0:b0 + chunk].clone())
        sets.append(bufs)
        steps.append(GraphedStep(bufs["teacher"], bufs["student"], bufs["theta"], bufs["flip"], bufs["dec"], bufs["w"], cfg,
                                 group=group, ema=None, instrument=False))
        filled.append(torch.cuda.Event())
        freed.append(torch.cuda.Event())
        freed[s].record(comp)

    def upload(i):
        s, b0 = i % 2, starts[i]
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(freed[s])                       # the step that read this set has finished
            sets[s]["teacher"].copy_(host["teacher"][:, :, b0:b0 + chunk], non_blocking=True)
            sets[s]["student"].copy_(host["student"][b0:b0 + chunk], non_blocking=True)
            sets[s]["theta"].copy_(host["theta"][:, b0:b0 + chunk], non_blocking=True)
            sets[s]["flip"].copy_(host["flip"][:, b0:b0 + chunk], non_blocking=True)
            sets[s]["dec"].copy_(dec[b0:b0 + chunk], non_blocking=True)
            sets[s]["w"].copy_(sample_w[b0:b0 + chunk], non_blocking=True)
            filled[s].record(copy_s)

    out = []
    upload(0)
    for i in range(len(starts)):
        s = i % 2
        if i + 1 < len(starts):
            upload(i + 1)                                     # overlaps the chain of chunk i
        comp.wait_event(filled[s])
        st = steps[s].run()
        if ema is not None and i == len(starts) - 1:
            ema.step(alpha)
        if on_chunk is not None:
            on_chunk(i, st)
        out.append(dict(summary=st["summary"].clone(), grad_scale=st["grad_scale"].clone(), count=st["count"].clone()))
        freed[s].record(comp)
    return out
