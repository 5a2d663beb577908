"""bench.py -- images/s of the SSD300 multibox post-backbone hot path (BASELINE.json metric).

A "step" is the post-backbone part of the reference's eval step (SSD_test_step,
SSD_trainer.py:214-256) over one batch of 256 synthetic images per GPU (BASELINE cfg 3):
    (i)  match + mined loss : build_targets + smooth-L1 + CELoss_w_neg_mining   (TR:92-117)
    (ii) decode + NMS       : predict(score 0.01, nms 0.45, max 200)            (SFS:388-476)
`value` = images / s over both halves with inputs resident in HBM; `parts` gives each half.
`e2e` is the same step through the drop-in Python API from pinned HOST buffers, with the
host->device copies of the head outputs / ground truth and the device->host read of the losses
and detections inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
N > 1: launched by torchrun, one rank per GPU; each rank owns 256 images (weak scaling) and the
only collective is the all-reduce of the three loss sums.
`--impl reference` times the UNMODIFIED reference (staged under oracle/_ref/ by oracle/stage_reference.py:
SSD_trainer.build_targets + smooth-L1 + CELoss_w_neg_mining, mySSD.predict) on the host cores for the same
metric (kind "reference"); the oracle port is timed beside it as a second figure.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "automotive-ssd-object-detection_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

BATCH = 256            # images per GPU (BASELINE.json metric: "at bs=256")
CFG = 3
N_SETS = 4             # rotating input sets: 4 x 89.5 MB > 126 MB L2, so no step re-reads a hot L2
P, C = 8732, 6
METRIC = "images/s for match+mined loss and decode+DIoU-NMS at bs=256 per GPU"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of each half's kernels at this workload, from the
# `ncu --set full` capture summarised in profiles/ (train_image_kernel + finalize_sums_kernel; predict_image_kernel)
WORKLOAD = ("cfg3: B=256/GPU eval-step post-backbone path on one (loc_all, conf_all) pair per batch, as SSD_test_step runs it "
            "(match+mined loss, then predict thr 0.01 / nms 0.45 / max 200), P=8732, C=6, G~U{1..20}")
TRAIN_KERNELS = "train_image_kernel + finalize_sums_kernel (match + encode + mined loss)"
PREDICT_KERNELS = "predict_image_kernel (stream -> row keys -> hot rows -> decode + rank + NMS, one launch)"
TRAFFIC = {"match_loss": 56.28e6 + 0.43e6 + 0.01e6, "decode_nms": 61.69e6 + 0.17e6, "source": "profiles/r02_final_ncu_full_summary.txt"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled while the timed region runs: NVML polled every
    millisecond (handle opened and queried once up front, so the first sample is not an NVML cold start);
    when NVML cannot be used, an `nvidia-smi -lms` child process (the B200_PROFILING.md clocks line)."""

    SMI_FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    SMI_NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag = index, period_s, [], set(), False
        self.max_mhz, self.nv, self.smi, self.err = None, None, None, None
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
        except Exception:
            uuid = None
        self.uuid = uuid
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode()) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception as e:                      # noqa: BLE001 -- any NVML failure selects the nvidia-smi path
            self.err = f"nvml: {type(e).__name__}: {e}"

    def start(self):
        if self.nv is None:
            import subprocess
            cmd = ["nvidia-smi", f"--query-gpu={self.SMI_FIELDS}", "--format=csv,noheader,nounits", "-lms", "10"]
            cmd += ["-i", self.uuid if self.uuid else str(self.index)]
            try:
                self.smi = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                time.sleep(0.3)                     # let the first lines arrive before the timed region starts
            except Exception as e:                  # noqa: BLE001
                self.err = f"{self.err}; nvidia-smi: {type(e).__name__}: {e}"
            self.t_start = time.time()
            return
        super().start()

    def run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while True:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:                  # noqa: BLE001
                self.err = f"nvml poll: {type(e).__name__}: {e}"
            if self.stop_flag:                      # checked after the query: a short region still gets one sample
                return
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        source = "nvml"
        if self.nv is not None:
            if self.ident is not None:
                self.join(timeout=2.0)
        elif self.smi is not None:
            source = "nvidia-smi -lms 10"
            time.sleep(0.05)
            self.smi.terminate()
            try:
                out, _ = self.smi.communicate(timeout=5)
            except Exception:
                out = ""
            rows = [r.split(",") for r in out.strip().splitlines() if r.count(",") >= 5]
            # lines printed before start() returned belong to the warm-up; keep the ones after it when there are any
            n_before = int(0.3 / 0.010)
            rows = rows[n_before:] if len(rows) > n_before else rows[-1:]
            for r in rows:
                try:
                    self.samples.append(int(float(r[0])))
                    self.max_mhz = int(float(r[1]))
                except ValueError:
                    continue
                for name, v in zip(self.SMI_NAMES, r[2:6]):
                    if v.strip().lower() == "active":
                        self.reasons.add(name)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["clock sampling unavailable"], "error": self.err}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": source}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm():
    """Which CPU implementation the baseline legs time: the UNMODIFIED reference staged under oracle/_ref/ (kind
    "reference": SSD_trainer.build_targets + SSD_trainer.py:104-108 + CELoss_w_neg_mining, and mySSD.predict on the
    precomputed head outputs), or -- only if that copy is missing -- the oracle port (kind "port")."""
    from oracle import refload as R
    if R.available():
        mdl = R.model("cpu")

        def train(cfg, n):
            R.train_half(cfg["loc_all"][:n], cfg["conf_infer"][:n], cfg["targets"][:n], cfg["iou_thresh"], cfg["ratio"], mdl)

        def pred(cfg, n):
            R.predict_half(cfg["loc_all"][:n], cfg["conf_infer"][:n], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"], False, mdl)
        return "reference", train, pred
    return ("port",) + port_arm()


def port_arm():
    from oracle import ssd_oracle as O
    pri, pri_xyxy = O.prior_tables()

    def train(cfg, n):
        O.train_half(pri, pri_xyxy, cfg["loc_all"][:n], cfg["conf_infer"][:n], cfg["targets"][:n], cfg["iou_thresh"], cfg["ratio"])

    def pred(cfg, n):
        O.postprocess(pri, cfg["loc_all"][:n], cfg["conf_infer"][:n], cfg["score_thresh"], cfg["nms_thresh"], cfg["max_per_img"], False)
    return train, pred


def cpu_reference_step(cfg, arm, n_train: int, n_pred: int):
    """One bounded sample of the workload on the CPU: returns (seconds train half, seconds predict
    half) for n_train / n_pred images.  The reference loops per image (TR:525, SFS:397), so its
    images/s does not depend on the batch size."""
    train, pred = arm
    t0 = time.perf_counter()
    with torch.no_grad():
        train(cfg, n_train)
        t1 = time.perf_counter()
        pred(cfg, n_pred)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_images_per_s(t_train, n_train, t_pred, n_pred):
    return 1.0 / (t_train / n_train + t_pred / n_pred)


CPU_KIND_NOTE = {"reference": "the unmodified reference (oracle/_ref: SSD_trainer.build_targets + smooth-L1 + CELoss_w_neg_mining; "
                              "mySSD.predict on precomputed head outputs)",
                 "port": "oracle port of the reference's eager per-image loop (oracle/_ref not staged)"}


def run_reference(args):
    from ssdhot import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_train, n_pred = 16, 2
    cfg = synth.config(CFG, batch=max(n_train, n_pred))
    kind, *arm = cpu_arm()
    for _ in range(max(args.warmup, 1) if args.warmup < 3 else 1):      # the CPU path has no clocks to warm; 1 pass pages code in
        cpu_reference_step(cfg, arm, 2, 1)
    # bound the run: size the per-step sample from one probe so K steps end within ~3 minutes
    probe = cpu_reference_step(cfg, arm, 2, 1)
    per_step_budget = 150.0 / max(args.steps, 1)
    n_pred = max(1, min(n_pred, int(per_step_budget * 0.7 / max(probe[1], 1e-3))))
    n_train = max(2, min(n_train, int(per_step_budget * 0.3 / max(probe[0] / 2, 1e-4))))
    tt = tp = 0.0
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        a, b = cpu_reference_step(cfg, arm, n_train, n_pred)
        tt += a
        tp += b
    wall = time.perf_counter() - t_begin
    value = cpu_images_per_s(tt, n_train * args.steps, tp, n_pred * args.steps)
    sample = (f"{n_train} images match+loss and {n_pred} images predict per step (per-image loop: images/s is batch independent); "
              + CPU_KIND_NOTE[kind])
    cpu = {"value": value, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample,
           "parts": {"match_loss_images_per_s": n_train * args.steps / tt, "decode_nms_images_per_s": n_pred * args.steps / tp}}
    if kind == "reference":                                              # the oracle port beside it, as a second figure
        parm = port_arm()
        cpu_reference_step(cfg, parm, 2, 1)
        a, b = cpu_reference_step(cfg, parm, n_train, n_pred)
        cpu["port"] = {"value": cpu_images_per_s(a, n_train, b, n_pred), "match_loss_images_per_s": n_train / a,
                       "decode_nms_images_per_s": n_pred / b, "note": CPU_KIND_NOTE["port"].split(" (")[0]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ssdhot", choices=["ssdhot", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of replaying a CUDA graph")
    ap.add_argument("--serial", action="store_true", help="run the two halves back to back on one stream instead of forked")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: all-reduce of the three loss sums by the peer-memory kernel (in the CUDA graph) or by NCCL")
    ap.add_argument("--collective-lag", type=int, default=0, choices=[0, 1],
                    help="peer collective: 0 = each step waits for its own reduced sums (what a training step needs; the headline), "
                         "1 = they arrive during the next step (eval-step logging only; also measured and printed as `lag1`)")
    ap.add_argument("--in-flight", type=int, default=2, choices=[1, 2, 3],
                    help="eval steps in flight (StepPipeline): consecutive eval batches are independent; 1 = strictly one after the other")
    ap.add_argument("--no-share", action="store_true", help="the two halves each stream conf_all (no key hand-off from the loss kernel to predict)")
    ap.add_argument("--quick", action="store_true", help="diagnostics: only the main timed loop (no halves, roofline, heads, e2e)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-large-batch", action="store_true", help="skip the extra B=2048 roofline measurement")
    ap.add_argument("--skip-strong", action="store_true", help="skip the strong-scaling cfg 4 section (B = 4096 over the job)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import ssdhot
    from ssdhot import dist as D, synth
    from ssdhot.engine import HotPathStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (ssdhot has no CPU fallback); use --impl reference for the CPU arm")
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    group = True if world > 1 else None

    # ---- synthetic inputs: N_SETS independent batches per rank, resident in HBM -------------------
    # one (loc_all, conf_all) pair per batch feeds both halves, as in SSD_test_step (SSD_trainer.py:208-256); conf = cfg 3's
    # inference logits (background biased, ~6.6 k candidates per image).  `conf_t` (no bias) only serves the `separate_conf`
    # variant: the round-1 workload, where the loss half read its own logits.
    spec = None
    sets = []
    host_sets = []
    for i in range(N_SETS):
        cfg = synth.config(CFG, batch=BATCH, seed_offset=rank * N_SETS + i)
        spec = spec or {k: cfg[k] for k in ("iou_thresh", "ratio", "score_thresh", "nms_thresh", "max_per_img")}
        host_sets.append(cfg)
        sets.append(dict(loc=cfg["loc_all"].to(dev), conf=cfg["conf_infer"].to(dev), conf_t=cfg["conf_train"].to(dev),
                         gt=ssdhot.pack_targets(cfg["targets"], dev)))
    ps = ssdhot.PriorSet.default(dev)
    # the sharded path's one exchange: every step's [sum loc, sum CE, sum positives] is all-reduced.  Default: the peer-memory
    # kernel (csrc/peer.cu) as a node of the step's CUDA graph, lag 0 (every step waits for its own reduced sums);
    # --collective nccl (or a box without CUDA IPC between the ranks): torch.distributed all-reduce on a side stream, outside
    # the graph (a captured NCCL collective hung at process exit on this stack)
    peer, peer_lag1, reducer, collective = None, None, None, "none"
    peers_more = []
    if group is not None:
        if args.collective == "peer":
            try:
                peer = D.PeerSums(dev, lag=args.collective_lag)
                peers_more = [D.PeerSums(dev, lag=args.collective_lag) for _ in range(args.in_flight - 1)]
                peer_lag1 = D.PeerSums(dev, lag=1) if args.collective_lag == 0 else None
                collective = ("peer-memory kernel over NVLink (ssdhot_allreduce_sums_peer), inside the step's CUDA graph, " +
                              ("lag 1: the reduced sums of step s are delivered during step s+1" if args.collective_lag else
                               "lag 0: every step waits for its own reduced sums"))
            except Exception as e:              # noqa: BLE001
                print(f"[bench] rank {rank}: PeerSums unavailable ({e}); using NCCL", file=sys.stderr)
        ok = torch.tensor([1 if (peer is not None or args.collective != "peer") else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if args.collective == "peer" and int(ok.item()) == 0:
            for pr in [peer, peer_lag1] + peers_more:
                if pr is not None:
                    pr.close()
            peer = peer_lag1 = None
            peers_more = []
        if peer is None:
            reducer = D.SumsReducer(dev)
            collective = "NCCL all-reduce (torch.distributed) on a side stream, outside the CUDA graph"

    def make_step(grp):
        return HotPathStep(ps, BATCH, C, spec["iou_thresh"], spec["ratio"], spec["score_thresh"], spec["nms_thresh"],
                           spec["max_per_img"], concurrent=not args.serial, group=grp)
    step = make_step(peer)
    step.share_keys = not args.no_share
    use_graph = not args.no_graph
    # the headline keeps --in-flight eval steps in flight (engine.StepPipeline): consecutive eval batches do not depend on each
    # other, so the loss kernel's logit stream of batch i+1 runs beside the NMS tail of batch i.  (The NCCL fallback reducer
    # works on one step at a time: depth 1 there.)
    from ssdhot.engine import StepPipeline
    depth = args.in_flight if (reducer is None and use_graph and not args.serial) else 1
    more_steps = [make_step(pr) for pr in (peers_more if group is not None else [None] * (depth - 1))][:depth - 1]
    for st_ in more_steps:
        st_.share_keys = step.share_keys
    pipe = StepPipeline(lambda k: ([step] + more_steps)[k], depth=depth)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(stp, n_steps, separate=False, use_reducer=True):
        """n_steps steps, CUDA events on the launching stream -> (ms total, max over ranks; per-rank list).  stp: a HotPathStep
        (one step after the other) or a StepPipeline (its steps in flight; the closing event waits for all of them)."""
        piped = isinstance(stp, StepPipeline)

        def one(i):
            s = sets[i % N_SETS]
            if piped:
                stp.submit(s["loc"], s["conf_t"] if separate else s["conf"], s["conf"], s["gt"], use_graph=use_graph)
                return
            stp.run(s["loc"], s["conf_t"] if separate else s["conf"], s["conf"], s["gt"], use_graph=use_graph)
            if reducer is not None and use_reducer:
                reducer.submit(stp.sums)
        # untimed: every (step object, input set) pair once, so that no CUDA-graph capture falls into the timed region however
        # small --warmup is; then the --warmup steps proper
        for so, st_ in (zip(stp.steps, stp.streams) if piped else [(stp, torch.cuda.current_stream(dev))]):
            with torch.cuda.stream(st_):
                for s in sets:
                    so.run(s["loc"], s["conf_t"] if separate else s["conf"], s["conf"], s["gt"], use_graph=use_graph)
        torch.cuda.synchronize(dev)
        for i in range(args.warmup):
            one(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            # device-side rendezvous: the ranks leave the host barrier milliseconds apart, and a rank whose timed region
            # started early would spend that skew waiting for its peers' first sums inside the region
            dist.all_reduce(torch.zeros((1,), device=dev))
        e0.record()
        for i in range(n_steps):
            one(i)
        if piped:
            stp.wait()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        per_rank = [float(ms.item())]
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            per_rank = [float(t.item()) for t in every]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), per_rank

    # launches per step (counted once, without the graph)
    s0 = sets[0]
    torch.cuda.synchronize(dev)
    n0 = ssdhot.launch_count()
    step.run(s0["loc"], s0["conf"], s0["conf"], s0["gt"], use_graph=False)
    launches_per_step = ssdhot.launch_count() - n0

    # ---- the headline: EXACTLY --steps steps, then a sustained region of >= 0.25 s under the same clock sampler ---------
    sampler = ClockSampler(local)           # (opens NVML: slow and uneven across ranks, so before the barrier)
    if not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    head = pipe if depth > 1 else step
    ms_total, per_rank_ms = timed_loop(head, args.steps)
    ms_per_step = ms_total / args.steps
    value = BATCH * world * args.steps / (ms_total / 1e3)
    n_long = max(args.steps, int(250.0 / max(ms_per_step, 1e-3)) + 1)
    ms_long, _ = timed_loop(head, n_long)
    clocks = sampler.result()
    count_picked = lambda: int((step.share[:4 * BATCH].view(torch.int32) == 3).sum().item()) if step.share_keys else 0
    picked_in_flight = count_picked()       # (the last step this object ran inside the headline's loop)
    ms_seq, _ = timed_loop(step, args.steps) if depth > 1 else (ms_total, None)
    sequential = {"steps_in_flight": 1, "ms_per_step": ms_seq / args.steps, "value": BATCH * world * args.steps / (ms_seq / 1e3),
                  "note": "the same steps strictly one after the other (one CUDA graph replay at a time on one stream): what rounds 1 and "
                          "2a reported as `value`, and what a TRAINING loop, whose steps depend on each other, can use"}
    sustained = {"steps": n_long, "ms_per_step": ms_long / n_long, "value": BATCH * world * n_long / (ms_long / 1e3),
                 "note": "the same loop run long enough (>= 0.25 s) for the clock sampler to cover it; `clocks` spans both regions"}
    # key hand-off (ssdhot.h: ssdhot_share_bytes): images whose predict CTA took the row keys the loss kernel's stream left
    picked = count_picked()
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "value": value, "ms_per_step": ms_per_step, "sustained": sustained,
                              "steps_in_flight": depth, "sequential": sequential,
                              "keys_handed_off": picked, "keys_handed_off_in_flight": picked_in_flight,
                              "per_rank_ms_per_step": [t / args.steps for t in per_rank_ms],
                              "collective": collective, "clocks": clocks}))
        if world > 1:
            dist.barrier()
            for pr in [peer, peer_lag1] + peers_more:
                if pr is not None:
                    pr.close()
            dist.destroy_process_group()
        return
    # the same step with each half streaming conf_all itself (the round-2 first-half form, for continuity)
    step.share_keys = False
    ms_ns, _ = timed_loop(step, args.steps)
    step.share_keys = not args.no_share
    handoff = {"on": step.share_keys, "images_handed_off_last_step": picked, "images_handed_off_last_step_in_flight": picked_in_flight,
               "batch": BATCH,
               "without": {"ms_per_step": ms_ns / args.steps, "value": BATCH * world * args.steps / (ms_ns / 1e3)},
               "note": "one conf_all feeds both halves (SSD_test_step): the loss kernel's logit stream leaves predict's 16-bit row keys "
                       "in HBM/L2 (17.5 KB per image) and predict_image_kernel picks them up instead of reading conf_all again; "
                       "`without` = both halves stream the logits"}
    # the round-1 workload (the loss half reads its own, unbiased logits: two conf tensors per step) for continuity
    ms_sep, _ = timed_loop(step, args.steps, separate=True)
    separate = {"ms_per_step": ms_sep / args.steps, "value": BATCH * world * args.steps / (ms_sep / 1e3),
                "note": "round-1 workload: conf_train for the loss half, conf_infer for predict (two logit tensors per step)"}
    lag1 = None
    if peer_lag1 is not None:
        step_l1 = make_step(peer_lag1)
        ms_l1, _ = timed_loop(step_l1, args.steps)
        lag1 = {"ms_per_step": ms_l1 / args.steps, "value": BATCH * world * args.steps / (ms_l1 / 1e3),
                "note": "peer collective with lag 1 (reduced sums arrive one step late: eval-step logging only, not a training step)"}

    # ---- per-half kernel time (CUDA events around each half, same rotation, direct launches) -------
    # (a step object without a group: with a PeerSums the loss forward leaves its final reduction to the exchange kernel,
    # and the halves are timed as a single GPU runs them -- loss kernel + finalize_sums_kernel -- at every N)
    step_sharded, step = step, (make_step(None) if group is not None else step)

    def time_half(train: bool, iters: int):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        st = torch.cuda.current_stream(dev).cuda_stream
        for i in range(3):
            s = sets[i % N_SETS]
            step.launch_loss(s["loc"], s["conf"], s["gt"], st) if train else step.launch_predict(s["loc"], s["conf"], st)
        torch.cuda.synchronize(dev)
        for i, (a, b) in enumerate(evs):
            s = sets[i % N_SETS]
            a.record()
            step.launch_loss(s["loc"], s["conf"], s["gt"], st) if train else step.launch_predict(s["loc"], s["conf"], st)
            b.record()
        torch.cuda.synchronize(dev)
        return statistics.mean(a.elapsed_time(b) for a, b in evs)

    iters = max(8, min(args.steps, 64))
    ms_loss = time_half(True, iters)
    ms_pred = time_half(False, iters)

    def time_half_two_in_flight(train: bool, n: int):
        """The same half with two launches in flight (two step objects on two streams, alternating input sets): elapsed / launches.
        Not a launch duration -- the rate at which the kernel retires batches when the 296 CTA slots stay filled (one launch of
        256 CTAs fills 0.86 of a wave and all its CTAs run their phases in lockstep)."""
        objs = [step, make_step(None)]
        strs = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        cur = torch.cuda.current_stream(dev)

        def go(i):
            s_, o_, q_ = sets[i % N_SETS], objs[i % 2], strs[i % 2]
            o_.launch_loss(s_["loc"], s_["conf"], s_["gt"], q_.cuda_stream) if train else o_.launch_predict(s_["loc"], s_["conf"], q_.cuda_stream)
        for q_ in strs:
            q_.wait_stream(cur)
        for i in range(6):
            go(i)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for q_ in strs:
            q_.wait_stream(cur)
        for i in range(n):
            go(i)
        for q_ in strs:
            cur.wait_stream(q_)
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    ms_loss2 = time_half_two_in_flight(True, 4 * iters)
    ms_pred2 = time_half_two_in_flight(False, 4 * iters)
    mean_g = statistics.mean(float(t["boxes"].shape[0]) for c in host_sets for t in c["targets"])
    k_mean = float(step.count.float().mean().item())
    bytes_loss = BATCH * (349296 + 24 * mean_g + 12)            # SURVEY.md 8d, path (i)
    bytes_pred = BATCH * (349296 + 28 * k_mean + 4)             # SURVEY.md 8d, path (ii)
    peak, peak_src = peaks()
    dom_is_pred = ms_pred >= ms_loss
    dom_ms, dom_bytes = (ms_pred, bytes_pred) if dom_is_pred else (ms_loss, bytes_loss)
    roofline = {
        "bound": "hbm", "kernel": PREDICT_KERNELS if dom_is_pred else TRAIN_KERNELS,
        "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": TRAFFIC["decode_nms" if dom_is_pred else "match_loss"],
        "traffic_source": TRAFFIC["source"], "peak_source": peak_src,
        # the same kernel's rate with two launches in flight, as the timed region of `value` keeps them (see parts_two_in_flight)
        "frac_two_in_flight": dom_bytes / ((ms_pred2 if dom_is_pred else ms_loss2) * 1e-3) / 1e9 / peak,
        "note": "algorithmic bytes = SURVEY.md 8(d) figure of the whole half (conf_all + loc_all + GT / outputs) over the CUDA-event "
                "time of that half run alone; reported for the slower half (both under `parts`)",
        "parts": {
            "match_loss": {"ms": ms_loss, "algorithmic_bytes": bytes_loss, "achieved": bytes_loss / (ms_loss * 1e-3) / 1e9,
                           "frac": bytes_loss / (ms_loss * 1e-3) / 1e9 / peak, "images_per_s": BATCH / (ms_loss * 1e-3)},
            "decode_nms": {"ms": ms_pred, "algorithmic_bytes": bytes_pred, "achieved": bytes_pred / (ms_pred * 1e-3) / 1e9,
                           "frac": bytes_pred / (ms_pred * 1e-3) / 1e9 / peak, "images_per_s": BATCH / (ms_pred * 1e-3)},
        },
        "parts_two_in_flight": {
            "note": "elapsed / launches with two launches of the half in flight on two streams (not a launch duration: the rate at which "
                    "the kernel retires batches once the 296 CTA slots stay filled; one launch of 256 CTAs is 0.86 of a wave)",
            "match_loss": {"ms_per_launch": ms_loss2, "achieved": bytes_loss / (ms_loss2 * 1e-3) / 1e9, "frac": bytes_loss / (ms_loss2 * 1e-3) / 1e9 / peak},
            "decode_nms": {"ms_per_launch": ms_pred2, "achieved": bytes_pred / (ms_pred2 * 1e-3) / 1e9, "frac": bytes_pred / (ms_pred2 * 1e-3) / 1e9 / peak},
        },
        # the whole forked step (what `value` times) against the same peak: as the two functions of the reference count their
        # bytes (each reads conf_all + loc_all), and as the step moves them once the loss kernel hands predict its row keys
        "step": {"ms": ms_per_step,
                 "algorithmic_bytes_two_functions": bytes_loss + bytes_pred,
                 "frac_two_functions": (bytes_loss + bytes_pred) / (ms_per_step * 1e-3) / 1e9 / peak,
                 "algorithmic_bytes_shared_inputs": bytes_loss + BATCH * (28 * k_mean + 4),
                 "frac_shared_inputs": (bytes_loss + BATCH * (28 * k_mean + 4)) / (ms_per_step * 1e-3) / 1e9 / peak},
    }

    # ---- the same halves fed by the head outputs: no permute / cat / pack pass ---------------
    def time_heads():
        from ssdhot.engine import HeadSet
        st_ = torch.cuda.current_stream(dev).cuda_stream
        out = {"note": "CUDA-event time of each half at this batch when the kernels read the six head outputs of each branch "
                       "directly (NCHW as the conv heads return them / channels_last), vs ssdhot_pack_heads + the packed kernels "
                       "and vs the reference's own permute + cat tail (SFS:249-269) + the packed kernels", "unit": "ms"}
        views = []
        for sset in sets[:3]:
            v = {}
            for name, cl in (("nchw", False), ("nhwc", True)):
                lh, ch = synth.heads_from_packed(sset["loc"], cl), synth.heads_from_packed(sset["conf"], cl)
                v[name] = dict(lh=lh, ch=ch, hs=HeadSet(lh, ch))
            views.append(v)

        def torch_tail(heads, D):
            return torch.cat([h.permute(0, 2, 3, 1).contiguous().view(h.shape[0], -1, D) for h in heads], 1)

        def run(mode, train, i):
            sset, v = sets[i % len(views)], views[i % len(views)]
            if mode in ("nchw", "nhwc"):
                step.launch_loss_heads(v[mode]["hs"], sset["gt"], st_) if train else step.launch_predict_heads(v[mode]["hs"], st_)
                return
            h = v["nchw"]
            if mode == "pack_then_packed":
                loc, conf = ssdhot.pack_heads(h["lh"], h["ch"])
            else:
                loc, conf = torch_tail(h["lh"], 4), torch_tail(h["ch"], C)
            step.launch_loss(loc, conf, sset["gt"], st_) if train else step.launch_predict(loc, conf, st_)

        for train in (True, False):
            res = {}
            for mode in ("nchw", "nhwc", "pack_then_packed", "torch_tail_then_packed"):
                for i in range(3):
                    run(mode, train, i)
                torch.cuda.synchronize(dev)
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(24)]
                for i, (a, b) in enumerate(evs):
                    a.record()
                    run(mode, train, i)
                    b.record()
                torch.cuda.synchronize(dev)
                res[mode] = statistics.median(a.elapsed_time(b) for a, b in evs)
            res["packed_kernels_only"] = ms_loss if train else ms_pred
            out["match_loss" if train else "decode_nms"] = res
        # the whole step (both halves forked, one CUDA graph per input set) from the heads, timed like `value`
        if group is None:
            whole = {}
            for mode in ("nchw", "nhwc"):
                for i in range(6):
                    step.run_heads(views[i % len(views)][mode]["hs"], views[i % len(views)][mode]["hs"], sets[i % len(views)]["gt"], use_graph=use_graph)
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = max(30, min(args.steps, 200))
                a.record()
                for i in range(n):
                    step.run_heads(views[i % len(views)][mode]["hs"], views[i % len(views)][mode]["hs"], sets[i % len(views)]["gt"], use_graph=use_graph)
                b.record()
                torch.cuda.synchronize(dev)
                whole[mode] = {"ms_per_step": a.elapsed_time(b) / n, "images_per_s": BATCH * n / (a.elapsed_time(b) * 1e-3)}
            out["whole_step"] = whole
        del views
        torch.cuda.empty_cache()
        return out

    heads = time_heads() if rank == 0 else None

    # ---- the HBM-bound streaming kernel of predict on its own (score_kernel: logits in, candidate lists out) -----
    def time_score_kernel(stp, inputs, batch, iters=16):
        """inputs: (loc, conf) pairs rotated over the launches (together larger than L2)."""
        st_ = torch.cuda.current_stream(dev).cuda_stream
        for i in range(3):
            stp.launch_predict(*inputs[i % len(inputs)], st_, stages=1)
        torch.cuda.synchronize(dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i, (a, b) in enumerate(evs):
            a.record(); stp.launch_predict(*inputs[i % len(inputs)], st_, stages=1); b.record()
        torch.cuda.synchronize(dev)
        ms_k = statistics.median(a.elapsed_time(b) for a, b in evs)
        n_cand = float(stp.pred_work[: batch * 16 * 4].view(torch.int32).sum().item())      # 16 list segments per image
        nbytes = batch * P * C * 4 + 8.0 * n_cand                                           # logits in, 8-byte keys out
        return {"per_gpu_batch": batch, "ms": ms_k, "algorithmic_bytes": nbytes, "achieved": nbytes / (ms_k * 1e-3) / 1e9,
                "frac": nbytes / (ms_k * 1e-3) / 1e9 / peak, "candidates_per_image": n_cand / batch}

    kernels = {"score_kernel": {"b256": time_score_kernel(step, [(x["loc"], x["conf"]) for x in sets], BATCH),
                                "note": "the generic two-kernel path's stream (logits in, 8-byte candidate keys out); the fast path is one kernel"}}
    roofline["kernels"] = kernels

    # ---- what a kernel that ONLY reads the logits costs at this batch (csrc/probe.cu): the practical floor under both halves ---
    def stream_floor(inputs, batch, iters=24):
        from ssdhot import _lib
        st_ = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty((batch,), dtype=torch.float32, device=dev)
        res = {}
        for mode, name in ((0, "ldg128"), (1, "bulk_async")):
            for i in range(3):
                _lib.check(_lib.lib().ssdhot_debug_stream_probe(inputs[i % len(inputs)].data_ptr(), batch, P * C * 4, mode, out.data_ptr(), st_), "probe")
            torch.cuda.synchronize(dev)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for i, (a, b) in enumerate(evs):
                a.record(); _lib.lib().ssdhot_debug_stream_probe(inputs[i % len(inputs)].data_ptr(), batch, P * C * 4, mode, out.data_ptr(), st_); b.record()
            torch.cuda.synchronize(dev)
            ms_k = statistics.median(a.elapsed_time(b) for a, b in evs)
            res[name] = {"ms": ms_k, "gbs": batch * P * C * 4 / (ms_k * 1e-3) / 1e9}
        return res
    roofline["stream_floor"] = {"b256": stream_floor([x["conf"] for x in sets], BATCH),
                                "note": "a kernel that only pulls conf_all (209,568 B per image, one CTA per image) and sums it, with the "
                                        "kernels' 16-byte loads or with cp.async.bulk: what the HBM side alone costs at this batch, launch included"}

    # ---- the same halves at B = 2048 per GPU (wave quantisation and launch latency amortised) -------------
    large = None
    if not args.skip_large_batch and rank == 0:
        LB = 2048
        cfg_l = synth.config(CFG, batch=LB, seed_offset=977)
        big = dict(loc=cfg_l["loc_all"].to(dev), conf=cfg_l["conf_infer"].to(dev), gt=ssdhot.pack_targets(cfg_l["targets"], dev))
        step_l = HotPathStep(ps, LB, C, spec["iou_thresh"], spec["ratio"], spec["score_thresh"], spec["nms_thresh"],
                             spec["max_per_img"])
        st = torch.cuda.current_stream(dev).cuda_stream

        def time_large(train: bool, iters: int = 12):
            f = (lambda: step_l.launch_loss(big["loc"], big["conf"], big["gt"], st)) if train else \
                (lambda: step_l.launch_predict(big["loc"], big["conf"], st))
            for _ in range(3):
                f()
            torch.cuda.synchronize(dev)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for a, b in evs:
                a.record(); f(); b.record()
            torch.cuda.synchronize(dev)
            return statistics.median(a.elapsed_time(b) for a, b in evs)

        ml, mp = time_large(True), time_large(False)
        kernels["score_kernel"]["b2048"] = time_score_kernel(step_l, [(big["loc"], big["conf"])], LB, iters=10)
        roofline["stream_floor"]["b2048"] = stream_floor([big["conf"]], LB, iters=10)
        g_l = statistics.mean(float(t["boxes"].shape[0]) for t in cfg_l["targets"])
        k_l = float(step_l.count.float().mean().item())
        bl, bp = LB * (349296 + 24 * g_l + 12), LB * (349296 + 28 * k_l + 4)
        large = {"per_gpu_batch": LB, "l2": "one input set of 0.7 GB (>> 126 MB L2)",
                 "match_loss": {"ms": ml, "achieved": bl / (ml * 1e-3) / 1e9, "frac": bl / (ml * 1e-3) / 1e9 / peak,
                                "images_per_s": LB / (ml * 1e-3)},
                 "decode_nms": {"ms": mp, "achieved": bp / (mp * 1e-3) / 1e9, "frac": bp / (mp * 1e-3) / 1e9 / peak,
                                "images_per_s": LB / (mp * 1e-3)}}
        del big, step_l, cfg_l
        torch.cuda.empty_cache()
    roofline["large_batch"] = large

    # ---- the layout-agnostic kernels (any priors / class count): cluster match + loss kernel, score + NMS kernels -------------
    generic = None
    if rank == 0:
        psg = ssdhot.PriorSet.default(dev, generic=True)
        step_g = HotPathStep(psg, BATCH, C, spec["iou_thresh"], spec["ratio"], spec["score_thresh"], spec["nms_thresh"], spec["max_per_img"])
        st_ = torch.cuda.current_stream(dev).cuda_stream

        def time_generic(fn, iters=24):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize(dev)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for i, (a, b) in enumerate(evs):
                a.record(); fn(i); b.record()
            torch.cuda.synchronize(dev)
            return statistics.median(a.elapsed_time(b) for a, b in evs)
        g_loss = time_generic(lambda i: step_g.launch_loss(sets[i % N_SETS]["loc"], sets[i % N_SETS]["conf"], sets[i % N_SETS]["gt"], st_))
        g_pred = time_generic(lambda i: (step_g.launch_predict(sets[i % N_SETS]["loc"], sets[i % N_SETS]["conf"], st_, stages=1),
                                         step_g.launch_predict(sets[i % N_SETS]["loc"], sets[i % N_SETS]["conf"], st_, stages=2)))
        generic = {"note": "the same halves on the layout-agnostic kernels (what other priors / class counts / > 64 boxes per image take): "
                           "gt_prepare + match_kernel (cluster of 8 CTAs per image) + loss_image_kernel + finalize; score_kernel + nms_image_kernel",
                   "match_loss_ms": g_loss, "decode_nms_ms": g_pred,
                   "match_loss_frac": bytes_loss / (g_loss * 1e-3) / 1e9 / peak, "decode_nms_frac": bytes_pred / (g_pred * 1e-3) / 1e9 / peak}
        del step_g
    roofline["generic_path"] = generic

    step = step_sharded
    barrier()       # (rank 0 alone measured the extra sections above; the sections below exchange sums again)
    # ---- strong scaling on BASELINE cfg 4: B = 4096 match + mined loss over the whole job, exchange inside the timed step ----
    strong = None
    if not args.skip_strong and 8 % world == 0:
        strong = strong_cfg4(ssdhot, synth, D, dist, ps, dev, rank, world, peer, spec, barrier)

    # ---- end to end through the drop-in API, host buffers ------------------------------------------
    # one step = ssdhot.eval_step on one (loc_all, conf_all) pair + the batch's ground truth as collate_detection packs it
    # (one pinned buffer); every step copies its own inputs host -> device and reads its own results back; the copies of
    # step i+1 are issued on a second stream before step i's kernels and read-back, so the PCIe transfer overlaps them
    # (double buffering, as a data loader would)
    pinned = []
    for cfg in host_sets[:2]:
        _, packed = ssdhot.collate_detection([(torch.zeros((1,)), t) for t in cfg["targets"]])
        pinned.append(dict(loc=cfg["loc_all"].pin_memory(), conf=cfg["conf_infer"].pin_memory(), conf_t=cfg["conf_train"].pin_memory(),
                           gt=packed.pin_memory()))
    copy_stream = torch.cuda.Stream(dev)
    d2h_holder = {}

    def h2d_probe():
        """A bare loop of the step's H2D copies (one cudaMemcpyAsync per buffer) on every rank at once: the ceiling the host
        side of e2e can reach on this box at this N (PCIe / NUMA), next to which e2e is read."""
        h = pinned[0]
        nbytes = sum(t.numel() * t.element_size() for t in (h["loc"], h["conf"], h["gt"].buffer))
        dst = [torch.empty_like(h["loc"], device=dev), torch.empty_like(h["conf"], device=dev), torch.empty_like(h["gt"].buffer, device=dev)]
        for _ in range(2):
            for d, t in zip(dst, (h["loc"], h["conf"], h["gt"].buffer)):
                d.copy_(t, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            for d, t in zip(dst, (h["loc"], h["conf"], h["gt"].buffer)):
                d.copy_(t, non_blocking=True)
        torch.cuda.synchronize(dev)
        gbs = torch.tensor([nbytes * n / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device=dev)
        every = [gbs.clone() for _ in range(world)]
        if world > 1:
            dist.all_gather(every, gbs)
        per_rank = [float(t.item()) for t in every]
        return {"bytes_per_step": nbytes, "per_rank_gbs": per_rank, "sum_gbs": sum(per_rank),
                "ceiling_images_per_s": BATCH * sum(per_rank) * 1e9 / nbytes,
                "note": "bare cudaMemcpyAsync loop of one step's inputs from pinned memory, all ranks at once"}

    def stage(i, separate):
        h = pinned[i % len(pinned)]
        with torch.cuda.stream(copy_stream):
            loc = h["loc"].to(dev, non_blocking=True)
            conf = h["conf"].to(dev, non_blocking=True)
            conf_t = h["conf_t"].to(dev, non_blocking=True) if separate else None
            gt = h["gt"].to(dev)                        # ONE copy: [offsets | boxes | labels]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return loc, conf, conf_t, gt, ev

    def e2e_compute(staged):
        loc, conf, conf_t, gt, ev = staged
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        for t in (loc, conf, conf_t):
            if t is not None:
                t.record_stream(cur)
        gt.record_stream(cur)
        grp = peer if peer is not None else group
        if conf_t is None:
            l_loc, l_conf, labels, scores, boxes, count = ssdhot.eval_step(ps, loc, conf, gt, spec["iou_thresh"], spec["ratio"],
                                                                           spec["score_thresh"], spec["nms_thresh"], spec["max_per_img"],
                                                                           group=grp)
        else:
            l_loc, l_conf = ssdhot.multibox_loss(ps, loc, conf_t, gt, spec["iou_thresh"], spec["ratio"], group=grp)
            labels, scores, boxes, count = ssdhot.predict_padded(ps, loc, conf, spec["score_thresh"], spec["nms_thresh"],
                                                                 spec["max_per_img"])
        out = (torch.stack((l_loc, l_conf)).cpu(), labels.cpu(), scores.cpu(), boxes.cpu(), count.cpu())
        d2h_holder["bytes"] = sum(t.numel() * t.element_size() for t in out)
        return out

    def e2e_run(n, separate):
        nxt = stage(0, separate)
        for i in range(n):
            cur_staged = nxt
            if i + 1 < n:
                nxt = stage(i + 1, separate)
            e2e_compute(cur_staged)

    def e2e_measure(separate):
        e2e_run(3, separate)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps, separate)
        torch.cuda.synchronize(dev)
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return BATCH * world * e2e_steps / float(t.item())

    probe = h2d_probe()
    e2e_steps = max(3, min(args.steps, 20))
    gt_bytes = pinned[0]["gt"].buffer.numel()
    h2d = BATCH * P * (4 + C) * 4 + gt_bytes
    e2e_value = e2e_measure(False)
    e2e_sep = e2e_measure(True)

    # ---- BASELINE cfg 1: one image, 5 boxes (the web app's case, app_files/ssd_demo_app.py:288): latency ------------
    single = None
    if rank == 0:
        c1 = synth.config(1)
        loc1, ci1 = c1["loc_all"].to(dev), c1["conf_infer"].to(dev)
        gt1 = ssdhot.pack_targets(c1["targets"], dev)
        step1 = HotPathStep(ps, 1, C, c1["iou_thresh"], c1["ratio"], c1["score_thresh"], c1["nms_thresh"], c1["max_per_img"])
        st1 = torch.cuda.current_stream(dev).cuda_stream

        def lat(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(dev)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for a, b in evs:
                a.record()
                fn()
                b.record()
            torch.cuda.synchronize(dev)
            return 1e3 * statistics.median(a.elapsed_time(b) for a, b in evs)

        def host_lat(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(n):
                t_a = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t_a)
            return 1e6 * statistics.median(ts)

        single = {
            "workload": "cfg1: B=1, 5 GT boxes, match+mined loss then predict(0.01, 0.45, 200)", "unit": "us",
            "kernels_match_loss": lat(lambda: step1.launch_loss(loc1, ci1, gt1, st1)),
            "kernels_decode_nms": lat(lambda: step1.launch_predict(loc1, ci1, st1)),
            "api_predict_list_of_dicts_wall": host_lat(lambda: ssdhot.predict(ps, None, c1["score_thresh"], c1["nms_thresh"], c1["max_per_img"],
                                                                                pre_loc_all=loc1, pre_conf_all=ci1)),
            "api_multibox_loss_item_wall": host_lat(lambda: [t.item() for t in ssdhot.multibox_loss(ps, loc1, ci1, gt1, c1["iou_thresh"], c1["ratio"])]),
        }

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        kind, *arm = cpu_arm()
        cfg = host_sets[0]
        cpu_reference_step(cfg, arm, 2, 1)
        n_train, n_pred = 256, 32          # ~0.5 s + ~8 s of CPU work on the 16-core box (bounded sample)
        a, b = cpu_reference_step(cfg, arm, n_train, n_pred)
        cpu = {"value": cpu_images_per_s(a, n_train, b, n_pred), "unit": "images/s", "cores": cores, "kind": kind,
               "sample": f"{n_train} images match+loss ({a:.2f} s) + {n_pred} images predict ({b:.2f} s), {CPU_KIND_NOTE[kind]}, "
                         "torch CPU threads = cores",
               "parts": {"match_loss_images_per_s": n_train / a, "decode_nms_images_per_s": n_pred / b}}
        if kind == "reference":
            parm = port_arm()
            cpu_reference_step(cfg, parm, 2, 1)
            a2, b2 = cpu_reference_step(cfg, parm, n_train, n_pred)
            cpu["port"] = {"value": cpu_images_per_s(a2, n_train, b2, n_pred), "match_loss_images_per_s": n_train / a2,
                           "decode_nms_images_per_s": n_pred / b2, "note": "oracle port of the same loop, same sample"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "global_batch": BATCH * world, "per_gpu_batch": BATCH,
                       "l2": f"inputs rotate over {N_SETS} sets of 89.5 MB (> 126 MB L2) so no step re-reads a warm L2",
                       "cuda_graph": use_graph, "halves": "serial" if args.serial else "forked (independent halves on two streams)",
                       "steps_in_flight": depth,
                       "steps_in_flight_note": "consecutive eval batches are independent (SSD_test_step): the timed region keeps this many steps "
                                               "in flight on as many streams (engine.StepPipeline), ms_per_step = elapsed / steps; `sequential` "
                                               "is the same run with one step at a time",
                       "parallelism": f"image-sharded x{world}, all-reduce of 3 doubles per step" + (f": {collective}" if world > 1 else "")},
            "parts": {"match_loss_images_per_s": BATCH * world / (ms_loss * 1e-3),
                      "decode_nms_images_per_s": BATCH * world / (ms_pred * 1e-3)},
            "sustained": sustained,
            "sequential": sequential,
            "separate_conf": separate,
            "key_handoff": handoff,
            "lag1": lag1,
            "strong_cfg4": strong,
            "roofline": roofline,
            "heads": heads,
            "single_image": single,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h_holder.get("bytes", 0), "steps": e2e_steps, "h2d_copies_per_step": 3,
                    "api": "ssdhot.eval_step (= SSD_test_step's post-backbone part: one loc_all, one conf_all, ground truth packed by "
                           "ssdhot.collate_detection into one pinned buffer) from pinned host tensors; the H2D copies of step i+1 overlap "
                           "the kernels and read-back of step i (two streams)",
                    "h2d_probe": probe,
                    "separate_conf": {"value": e2e_sep, "h2d_bytes_per_step": h2d + BATCH * P * C * 4, "h2d_copies_per_step": 4,
                                      "note": "round-1 workload: ssdhot.multibox_loss(conf_train) + ssdhot.predict_padded(conf_infer)"}},
            "gpu_launches": launches_per_step * args.steps,
            "per_rank_ms_per_step": [t / args.steps for t in per_rank_ms],
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        for pr in [peer, peer_lag1] + peers_more:
            if pr is not None and pr.timed_out():
                print(f"[bench] rank {rank}: a peer all-reduce timed out", file=sys.stderr)
        dist.barrier()
        for pr in [peer, peer_lag1] + peers_more:
            if pr is not None:
                pr.close()
        dist.destroy_process_group()


def strong_cfg4(ssdhot, synth, D, dist, ps, dev, rank, world, peer, spec, barrier):
    """BASELINE cfg 4: one batch of 4096 images (G ~ U{1..20}), match + mined loss only, split over the job's ranks in
    contiguous image slices, the all-reduce of the three sums INSIDE the timed step (lag 0).  The batch is generated as 8
    blocks of 512 images (seeds independent of N) so that every N in {1, 2, 4, 8} sees the same 4096 images; rank r owns
    blocks [r * 8 / N, (r + 1) * 8 / N).  `sharded_check`: rank 0 also runs the WHOLE batch alone and compares its sums with
    the reduced sums of the sharded run (the logic of tools/check_sharded_nccl.py, now part of the driver-run record)."""
    from ssdhot.engine import HotPathStep
    TOTAL, BLOCK = 4096, 512
    per = 8 // world
    mine = range(rank * per, (rank + 1) * per)

    def block(k):
        return synth.config(4, batch=BLOCK, seed_offset=500 + k)

    def to_dev(cfgs):
        loc = torch.cat([c["loc_all"] for c in cfgs], 0).to(dev)
        conf = torch.cat([c["conf_infer"] for c in cfgs], 0).to(dev)
        gt = ssdhot.pack_targets([t for c in cfgs for t in c["targets"]], dev)
        return loc, conf, gt

    my_cfgs = [block(k) for k in mine]
    loc, conf, gt = to_dev(my_cfgs)
    n_local = loc.shape[0]
    grp = None
    if world > 1:
        grp = peer if peer is not None and peer.lag == 0 else (D.PeerSums(dev, lag=0) if peer is not None else True)
    stp = HotPathStep(ps, n_local, C, spec["iou_thresh"], spec["ratio"], infer_half=False, group=grp)
    st = torch.cuda.current_stream(dev)

    def one():
        stp.launch_loss(loc, conf, gt, st.cuda_stream)
        stp._reduce()

    for _ in range(5):
        one()
    barrier()
    if world > 1:
        dist.all_reduce(torch.zeros((1,), device=dev))
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        one()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sharded_sums = stp.sums.clone()
    out = {"workload": "cfg4: B=4096, G~U{1..20}, match + mined loss, all-reduce of the 3 sums inside the timed step (lag 0)",
           "scaling": "strong", "global_batch": TOTAL, "per_gpu_batch": n_local, "ms_per_step": float(ms.item()),
           "images_per_s": TOTAL / (float(ms.item()) * 1e-3),
           "l2": f"one input set of {n_local * 349296 / 1e6:.0f} MB per rank" + (" (> 126 MB L2)" if n_local * 349296 > 126e6 else " (fits the 126 MB L2: re-read warm)"),
           "note": "speed-up over N = 1 = this line's images_per_s / the N = 1 run's (the driver has both lines)"}
    check = None
    if rank == 0:
        if world > 1:
            del loc, conf, gt
            rest = [block(k) for k in range(8) if k not in mine]
            order = sorted(list(mine) + [k for k in range(8) if k not in mine])
            cfgs = {k: c for k, c in zip(list(mine) + [k for k in range(8) if k not in mine], my_cfgs + rest)}
            f_loc, f_conf, f_gt = to_dev([cfgs[k] for k in order])
            whole = HotPathStep(ps, TOTAL, C, spec["iou_thresh"], spec["ratio"], infer_half=False)
            whole.launch_loss(f_loc, f_conf, f_gt, st.cuda_stream)
            torch.cuda.synchronize(dev)
            single_sums = whole.sums.clone()
        else:
            single_sums = sharded_sums
        a, b = sharded_sums.cpu(), single_sums.cpu()
        rel = float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())
        check = {"ok": bool(rel <= 1e-9 and a[2].item() == b[2].item()), "max_rel_err": rel,
                 "sums_sharded": a.tolist(), "sums_single_gpu": b.tolist(),
                 "what": "[sum smooth-L1, sum CE, sum positives] of the 4096-image batch: reduced over the ranks vs one GPU alone"}
    out["sharded_check"] = check
    if world > 1 and isinstance(grp, D.PeerSums) and grp is not peer:
        barrier()
        grp.close()
    barrier()
    return out


if __name__ == "__main__":
    main()
