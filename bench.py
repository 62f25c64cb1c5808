#!/usr/bin/env python
"""bench.py -- depth-maps/sec of the CER-MVS inference hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the HOT PATH ONLY (core/raft.py:75-108: 2 cost-volume builds + 16+16 GRU iterations; the
encoders fnet / cnet of core/extractor.py are NOT part of it, their feature / context maps are the inputs) over one
synthetic reference image of the chosen BASELINE.json configuration:

    cfg2 (default)  DTU 1184x1600, 10 source views                     -- the configuration the metric is quoted on
    cfg3            DTU cascaded two-pass: cfg2 + the rescale=2 pass (2368x3200, 592x800 grid) + disp -> depth +
                    multires merge (multires.py:24-28) per step
    cfg4            Tanks&Temples 1056x1920, 15 source views           -- at N > 1 ONE image sharded over the N GPUs
    cfg5            BlendedMVS 1536x2048, 7 source views               -- at N > 1 ONE image sharded over the N GPUs

One process per GPU.  cfg2 / cfg3 at N > 1: every rank works on its own reference image (replicas, weak scaling, no
data-path collective); the sharded single-image path (every rank builds its run of (view, hypothesis) units, ONE NCCL
all-reduce of the partial cost volume per stage) is timed next to it at every N.  cfg4 / cfg5 at N > 1: the sharded
path IS the headline (strong scaling), as BASELINE.json words those configurations.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own Python (baseline/_ref, imported
unmodified; its CUDA-only alt_cuda_corr.forward served by the oracle's CPU restatement) on all host threads on a
bounded sample of the same workload.  The reference's GPU path on the same B200 is timed in the `reference_gpu` block
of our own line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from cer_mvs_b200 import synth  # noqa: E402

CASCADE = [(64, 64, 16), (-1, 320, 16)]   # "32 iters": 16 + 16 (reference default is 8 + 8, core/raft.py:16)
CASCADE_8 = [(64, 64, 8), (-1, 320, 8)]
CONFIGS = {
    # name: (H, W, V, sharded headline at N > 1, description)
    "cfg2": (1184, 1600, 10, False, "DTU 1184x1600 (296x400 grid), 10 source views (BASELINE configs[1])"),
    "cfg3": (1184, 1600, 10, False, "DTU cascaded two-pass: 1184x1600 + rescale=2 pass 2368x3200 (592x800 grid), 10 source "
                                    "views, disp->depth + multires merge per step (BASELINE configs[2])"),
    "cfg4": (1056, 1920, 15, True, "Tanks&Temples 1056x1920 (264x480 grid), 15 source views (BASELINE configs[3])"),
    "cfg5": (1536, 2048, 7, True, "BlendedMVS 1536x2048 (384x512 grid), 7 source views (BASELINE configs[4])"),
}
H, W, V = CONFIGS["cfg2"][:3]
SCOPE = "HOT PATH ONLY (cost-volume builds + GRU lookup/update iterations, core/raft.py:75-108); encoders excluded, their " \
        "fp16 feature / context maps are the inputs"
METRIC = "depth-maps/sec (DTU 1600x1184, 10 src views, 32 iters)"


def workload_name(cfg):
    return f"{CONFIGS[cfg][4]}, 16+16 GRU iterations, fp16 features; {SCOPE}"


WORKLOAD = workload_name("cfg2")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own Python on the host cores (baseline/bench_ref.py)
# ---------------------------------------------------------------------------------------------
def host_threads():
    """All host threads the process may use; set explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def cpu_sample(rows=24, iters=(1, 1)):
    """One bounded sample of the reference on the host: (seconds per full depth map, description, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import bench_ref
    full, desc, _ = bench_ref.reference_cpu(H, W, V, CASCADE, rows, iters)
    return full, desc, "reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import bench_ref
    for _ in range(args.warmup):
        cpu_sample(rows=args.ref_rows, iters=(1, 1))
    times = []
    desc = kind = ""
    t0 = time.perf_counter()
    for _ in range(args.steps):
        full, desc, kind = cpu_sample(rows=args.ref_rows, iters=(1, 1))
        times.append(full)
    wall = time.perf_counter() - t0
    sec = float(np.mean(times))
    val = 1.0 / sec
    cfg1_s = bench_ref.reference_cpu_cfg1()          # one whole, un-extrapolated depth map of BASELINE configs[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "depth-maps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": "host cores", "l2": "n/a (CPU)",
                   "engine": "reference Python (baseline/_ref), hot path only"},
        "cpu_baseline": {"value": val, "unit": "depth-maps/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "depth-maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cfg1_whole_depth_map": {"seconds": cfg1_s, "depth_maps_per_s": 1.0 / cfg1_s, "extrapolated": False,
                                 "workload": "BASELINE configs[0]: 448x576, 2 source views, 2+2 iterations, fp32"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        # the sampler runs from before the warm-up; keep the samples that arrived inside the timed region
        rows = [r for (ts, r) in self.rows if self.t0 is None or (self.t0 <= ts <= (self.t1 or ts) + 0.05)]
        if not rows:
            rows = [r for (_, r) in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
def algorithmic(px, n_views):
    """Algorithmic work per launch of each kernel class."""
    return {
        "conv_gates": ("tensor", 2.0 * px * 9 * 64 * (241 + 241 + 177)),
        "conv_q_gru": ("tensor", 2.0 * px * 9 * 64 * 64),
        "conv_delta": ("tensor", 2.0 * px * (9 * 64 * 256 + 9 * 256)),
        "conv_corr_enc_3x3": ("tensor", 2.0 * px * 9 * 64 * 64),
        "corr_enc_1x1": ("tensor", 2.0 * px * 33 * 64),
        # lookup: disp + origin + 3 windows of 12 floats read, 33 floats written = 284 B / pixel
        "lookup": ("hbm", 284.0 * px),
        # fused build per stage: (V+1) feature maps fp16 + disp + volume write; D averaged over the two stages
        "volume_build": ("hbm", (n_views + 1) * px * 64 * 2.0 + 4.0 * px + 4.0 * px * (64 + 44) / 2),
    }


class Scene:
    """Pinned host inputs + device copies of one synthetic reference image."""

    def __init__(self, Hh, Ww, Vv, seed, dev):
        t = torch.from_numpy
        sc = synth.make_scene(Hh, Ww, Vv, seed=seed)
        self.H, self.W, self.V = Hh, Ww, Vv
        self.h1, self.w1 = Hh // 4, Ww // 4
        self.np_poses, self.np_K = sc["poses"], sc["intrinsics"]
        self.h_fm = t(sc["fmaps"]).half().pin_memory()
        self.h_net = t(sc["net"]).half().pin_memory()
        self.h_inp = t(sc["inp"]).half().pin_memory()
        self.d_fm, self.d_net, self.d_inp = self.h_fm.to(dev), self.h_net.to(dev), self.h_inp.to(dev)
        self.d_poses, self.d_K = t(sc["poses"]).to(dev), t(sc["intrinsics"]).to(dev)
        self.sc = sc

    @property
    def h2d_bytes(self):
        return (self.h_fm.numel() + self.h_net.numel() + self.h_inp.numel()) * 2 + (self.V + 1) * (16 + 9) * 4

    def dev_args(self):
        return (self.d_fm, self.d_net, self.d_inp, self.d_poses, self.d_K, 1.0)

    def host_args(self):
        return (self.h_fm, self.h_net, self.h_inp, self.np_poses, self.np_K, 1.0)


def run_ours(args):
    import torch.distributed as dist
    from cer_mvs_b200 import _lib, prep
    from cer_mvs_b200.hotpath import DepthHotPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (cer_mvs_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().cer_device_check(), "device check")
    if args.conv_variant is not None:
        _lib.check(_lib.lib().cer_set_conv_variant(args.conv_variant), "conv variant")
    if args.build_variant is not None:
        _lib.check(_lib.lib().cer_set_build_variant(args.build_variant), "build variant")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cfg = args.config
    Hc, Wc, Vc, sharded_headline, _ = CONFIGS[cfg]
    sharded_headline = sharded_headline and world > 1
    two_pass = cfg == "cfg3"
    h1, w1 = Hc // 4, Wc // 4
    px = h1 * w1
    sd = synth.make_update_weights(seed=0, delta_scale=0.1, delta_bias=0.005)

    # replicas: every rank its own image; sharded headline: every rank the same image
    scene = Scene(Hc, Wc, Vc, 0 if sharded_headline else rank, dev)
    hp = DepthHotPath(h1, w1, max_views=Vc, cascade=CASCADE, feats_f16=True, use_graph=not args.profile_step)
    hp.load_update_block(sd)
    scene2 = hp2 = None
    if two_pass:                      # the rescale=2 pass: an independent full inference at twice the image size
        scene2 = Scene(2 * Hc, 2 * Wc, Vc, rank, dev)
        hp2 = DepthHotPath(2 * h1, 2 * w1, max_views=Vc, cascade=CASCADE, feats_f16=True)
        hp2.load_update_block(sd)

    def step_device():
        if sharded_headline:
            return hp.forward_sharded(*scene.dev_args())
        out = hp(*scene.dev_args())
        if two_pass:                  # demo.py:27-43: inference(rescale=1), inference(rescale=2), multires()
            out2 = hp2(*scene2.dev_args())
            d1 = prep.disp_to_depth(out.view(h1, w1))
            d2 = prep.disp_to_depth(out2.view(2 * h1, 2 * w1))
            return prep.multires_merge(d1, d2, 0.02)
        return out

    if args.profile_step:       # for ncu: W warm-up steps, then exactly one eager step, nothing else
        for _ in range(args.warmup + 1):
            step_device()
        torch.cuda.synchronize()
        print(json.dumps({"profile_step": True, "launches_per_step": hp.last_launch_count}))
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- device-resident throughput (value) ----
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    clocks.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    clocks.mark_end()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    launches_per_step = hp.last_launch_count + (hp2.last_launch_count + 3 if two_pass else 0)
    launches = launches_per_step * args.steps
    jobs = 1 if sharded_headline else world          # depth maps finished per step over all ranks
    value = jobs * args.steps / (ms / 1e3)

    # ---- the reference's default iteration count, 8 + 8 (core/raft.py:16), device-resident, for SURVEY 8d ----
    ms8 = None
    if cfg == "cfg2":
        hp8 = DepthHotPath(h1, w1, max_views=Vc, cascade=CASCADE_8, feats_f16=True)
        hp8.load_update_block(sd)
        n8 = max(args.steps // 2, 3)
        ms8 = timed(lambda: hp8(*scene.dev_args()), n8, args.warmup) / n8
        del hp8

    # ---- end to end with host buffers (e2e): every step copies its inputs from pinned host memory and reads its
    # result back ----
    h2d = scene.h2d_bytes + (scene2.h2d_bytes if two_pass else 0)
    d2h = (4 * px if two_pass else px) * 4
    sync_ms = None
    if not sharded_headline and not two_pass:
        # the copies of step i+1 overlap the kernels of step i (cer_plan_submit_host, two jobs in flight)
        h_outs = [torch.empty(1, 1, h1, w1).pin_memory() for _ in range(2)]
        for i in range(min(args.warmup, 2)):
            hp.submit_host(*scene.host_args(), out=h_outs[i & 1])
        hp.wait_host()
        hp.wait_host()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            hp.submit_host(*scene.host_args(), out=h_outs[i & 1])
            if i >= 1:
                hp.wait_host()            # result of step i-1 is on the host
        hp.wait_host()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        t0 = time.perf_counter()
        h_out = torch.empty(1, 1, h1, w1).pin_memory()
        for _ in range(3):               # latency of one synchronous call (copies not overlapped), for reference
            hp.run_host(*scene.host_args(), out=h_out)
        sync_ms = 1e3 * (time.perf_counter() - t0) / 3
        e2e_api = "cer_plan_submit_host / cer_plan_wait_host (pinned host buffers, 2 jobs in flight)"
    else:
        # sharded / two-pass: explicit pinned -> device copies on the stream, the public device API, result back to a
        # pinned host buffer; one synchronisation per step
        scenes = [scene] + ([scene2] if two_pass else [])
        h_res = torch.empty((2 * h1, 2 * w1) if two_pass else (1, 1, h1, w1)).pin_memory()

        def step_host():
            for sc_ in scenes:
                sc_.d_fm.copy_(sc_.h_fm, non_blocking=True)
                sc_.d_net.copy_(sc_.h_net, non_blocking=True)
                sc_.d_inp.copy_(sc_.h_inp, non_blocking=True)
                sc_.d_poses.copy_(torch.from_numpy(sc_.np_poses), non_blocking=False)
                sc_.d_K.copy_(torch.from_numpy(sc_.np_K), non_blocking=False)
            out = step_device()
            h_res.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_api = ("DepthHotPath.forward_sharded" if sharded_headline else "DepthHotPath x2 + prep.multires_merge") + \
                  " on device tensors; pinned host -> device copies and the result read-back inside every step"
    e2e = jobs * args.steps / e2e_s

    # ---- per-kernel breakdown with CUDA events (eager), roofline of the dominant kernel ----
    hp.set_kernel_timing(True)
    hp(*scene.dev_args())
    hp.kernel_times()
    ksteps = min(args.steps, 5)
    for _ in range(ksteps):
        hp(*scene.dev_args())
    kt = hp.kernel_times()
    hp.set_kernel_timing(False)
    pk = peaks()
    alg = algorithmic(px, Vc)
    kernels = {}
    for k, (tot, n) in kt.items():
        if n == 0:
            continue
        if k == "volume_build":       # one build per cascade stage = the staged kernel + the gather pass for flagged tiles
            n = 2 * ksteps
        ent = {"ms_per_step": tot / ksteps, "launches_per_step": n / ksteps, "avg_us": 1e3 * tot / n}
        if k in alg:
            bound, work = alg[k]
            ach = work / (tot / n * 1e-3) / (1e9 if bound == "hbm" else 1e12)
            ent.update(bound=bound, achieved=ach, frac=ach / pk[bound], unit="GB/s" if bound == "hbm" else "TFLOP/s")
        kernels[k] = ent
    ksum = sum(v["ms_per_step"] for v in kernels.values())
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    tjson = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from ncu --set full captures
    if os.path.isfile(tp):
        tjson = json.load(open(tp))

    def roof_of(k):
        return {"kernel": k, "bound": kernels[k].get("bound"), "achieved": kernels[k].get("achieved"),
                "peak": pk.get(kernels[k].get("bound", "tensor")), "unit": kernels[k].get("unit"),
                "frac": kernels[k].get("frac"), "traffic": tjson.get(k) if cfg == "cfg2" else None,
                "peak_source": pk["src"], "avg_launch_us": kernels[k]["avg_us"],
                "share_of_step": kernels[k]["ms_per_step"] / ksum,
                "share_note": "share of the SUM of event-timed kernels of an eager pass (an event pair adds ~6 us to "
                              "every short launch, so the sum exceeds the graph-replayed ms_per_step)"}
    roof = roof_of(dom)
    build_roof = roof_of("volume_build") if "volume_build" in kernels else None
    if build_roof is not None:
        build_roof["note"] = ("algorithmic HBM bytes; the kernel's real bound is arithmetic: 512 flop per sample "
                              "(px * D * V samples) = %.1f GFLOP per launch on average" %
                              (512.0 * px * Vc * 54 / 1e9))
    lookup_plan_roof = roof_of("lookup") if "lookup" in kernels else None

    # ---- the stand-alone correlation-lookup kernel (CorrBlock.__call__ drop-in, cer_lookup) on per-view volumes:
    # slots = V, so one launch moves V * px * (4 D + 8 + 132) bytes (larger than L2, nothing is reused between
    # launches); algorithmic bytes per SURVEY 8d = slots * px * 284 ----
    lookup_roof = None
    if rank == 0 and cfg == "cfg2":
        L = _lib.lib()
        Dl, incre = 64, 0.0025 / 64
        g = torch.Generator(device=dev).manual_seed(1)
        vol = torch.randn(Vc, px, Dl, device=dev, generator=g)
        origin = torch.full((px,), 32 * incre, device=dev)
        zinv = origin + (torch.rand(px, device=dev, generator=g) * 40 - 20) * incre
        lout = torch.empty(Vc, 33, px, device=dev)
        st = _lib.stream_ptr()

        def run_lookup():
            _lib.check(L.cer_lookup(vol.data_ptr(), Vc, origin.data_ptr(), zinv.data_ptr(), Dl, incre, 5, 3,
                                    lout.data_ptr(), h1, w1, st), "cer_lookup")
        for _ in range(3):
            run_lookup()
        torch.cuda.synchronize()
        n_l = 20
        e0.record()
        for _ in range(n_l):
            run_lookup()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / n_l
        alg_b, moved_b = Vc * px * 284.0, Vc * px * (4.0 * Dl + 8 + 132)
        lookup_roof = {"kernel": "lookup_v2_kernel<64> (cer_lookup, slots = V per-view volumes)", "bound": "hbm",
                       "achieved": alg_b / (us * 1e-6) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                       "frac": alg_b / (us * 1e-6) / 1e9 / pk["hbm"], "avg_launch_us": us,
                       "design_bytes_per_launch": moved_b, "design_gbs": moved_b / (us * 1e-6) / 1e9,
                       "design_frac": moved_b / (us * 1e-6) / 1e9 / pk["hbm"],
                       "note": "rows are read whole (4D+8 B in, 132 B out per pixel-slot = 396 B at D=64), so 284 "
                               "algorithmic bytes cap 'frac' at 0.72 of the achieved HBM fraction ('design_frac')",
                       "traffic": tjson.get("lookup_dropin_slotsV"),
                       "in_plan_kernel": lookup_plan_roof}
        del vol, lout

    # ---- the build kernels on their own (cer_build_volume, same features / cameras): stage 0 as in the step, and stage 1
    # on the scene's TRUE (smooth) disparity map -- what a trained model hands to the second stage.  The step above runs
    # random-init GRU weights, whose stage-1 input is noise of ~60 hypothesis steps between neighbouring pixels: every
    # 16x8 tile is incoherent there and takes the gather pass, so the staged kernel's stage-1 time never shows in it. ----
    if rank == 0 and cfg == "cfg2" and build_roof is not None:
        L = _lib.lib()
        st = _lib.stream_ptr()
        feats = torch.empty(Vc + 1, h1, w1, 64, device=dev, dtype=torch.float16)
        _lib.check(L.cer_nchw_to_nhwc(scene.d_fm.data_ptr(), 1, feats.data_ptr(), 1, Vc + 1, 64, h1, w1, 0.125, st))
        ii = torch.zeros(Vc, dtype=torch.int32, device=dev)
        jj = torch.arange(1, Vc + 1, dtype=torch.int32, device=dev)
        Kq = scene.d_K[0].clone()
        Kq[:, :2] /= 4
        Pq = scene.d_poses[0].contiguous()
        Pij = torch.empty(Vc, 16, device=dev)
        _lib.check(L.cer_projection_matrices(Pq.data_ptr(), Kq.contiguous().data_ptr(), ii.data_ptr(), jj.data_ptr(), Vc,
                                             Pij.data_ptr(), st))
        origin = torch.zeros(h1, w1, device=dev)
        disp_in = [torch.zeros(h1, w1, device=dev), torch.from_numpy(scene.sc["true_disp"]).to(dev).contiguous()]
        stages = [(64, 0.0025 / 64, 1), (44, 0.0025 / 320, 0)]
        coh = {"what": "cer_build_volume alone, event-timed: stage 0 (zero disparity + shift) and stage 1 on the scene's true "
                       "(smooth) disparity map; 'staged' = default (TMA-staged source boxes + tcgen05, gather pass for "
                       "incoherent tiles), 'gather' = cer_set_build_variant(1)"}
        try:
            for variant, name in ((0, "staged"), (1, "gather")):
                _lib.check(L.cer_set_build_variant(variant))
                for sidx, (Dd, inc, shift) in enumerate(stages):
                    vol = torch.empty(px, Dd, device=dev)
                    lo = float(torch.tensor(Dd // 2 * inc).float())

                    def run_build():
                        _lib.check(L.cer_build_volume(feats.data_ptr(), 1, Pij.data_ptr(), ii.data_ptr(), jj.data_ptr(), Vc,
                                                      disp_in[sidx].data_ptr(), shift, Dd, inc, lo, origin.data_ptr(),
                                                      vol.data_ptr(), 1.0 / Vc, 0, h1, w1, st), "cer_build_volume")
                    for _ in range(3):
                        run_build()
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(10):
                        run_build()
                    e1.record()
                    torch.cuda.synchronize()
                    ms_b = e0.elapsed_time(e1) / 10
                    alg_b = (Vc + 1) * px * 128.0 + 4.0 * px + 4.0 * px * Dd
                    coh[f"{name}_stage{sidx}_ms"] = ms_b
                    coh[f"{name}_stage{sidx}_frac"] = alg_b / (ms_b * 1e-3) / 1e9 / pk["hbm"]
        finally:
            L.cer_set_build_variant(args.build_variant if args.build_variant is not None else 0)
        build_roof["coherent"] = coh
        del feats

    # ---- one image sharded over all ranks (NCCL all-reduce of the partial volume per stage), at every N > 1 ----
    sharded = None
    if world > 1 and not sharded_headline and not two_pass:
        sc0 = scene if rank == 0 else Scene(Hc, Wc, Vc, 0, dev)
        vms = timed(lambda: hp.forward_sharded(*sc0.dev_args()), args.steps, 3) / args.steps
        sharded = {"ms_per_depth_map": vms, "depth_maps_per_s": 1e3 / vms,
                   "units": f"{Vc} views x (64 | 44) hypotheses split evenly over {world} ranks",
                   "collective": "nccl all_reduce(sum) of the partial volume, "
                                 f"{px * 64 * 4 / 1e6:.1f}+{px * 44 * 4 / 1e6:.1f} MB per depth map"}
    elif sharded_headline:
        # ... and the replica mode next to the sharded headline
        rms = timed(lambda: hp(*scene.dev_args()), args.steps, 3) / args.steps
        sharded = {"replicas_depth_maps_per_s": world * 1e3 / rms, "replicas_ms_per_depth_map": rms}

    # ---- whole RAFT.forward: images in -> encoders (csrc/encoder.cu) -> hot path, all on our kernels ----
    whole = None
    if rank == 0 and world == 1 and cfg == "cfg2" and not args.no_whole_forward:
        from cer_mvs_b200.raft import RAFT
        sdw = {}
        for k, v in synth.make_encoder_weights(seed=0, out_dim=64).items():
            sdw["fnet." + k] = torch.from_numpy(v)
        for k, v in synth.make_encoder_weights(seed=1, out_dim=128).items():
            sdw["cnet." + k] = torch.from_numpy(v)
        for k, v in sd.items():
            sdw["update_block." + k] = torch.from_numpy(v)
        model = RAFT(cascade=CASCADE, test_mode=True)
        model.load_state_dict(sdw, strict=True)
        model = model.to(dev).eval()
        h_images = torch.from_numpy(synth.make_image(Hc, Wc, n=Vc + 1, seed=0))[None].pin_memory()
        d_images = h_images.to(dev)
        with torch.no_grad():
            wms = timed(lambda: model(d_images, scene.d_poses, scene.d_K, scale=1.0), max(args.steps // 2, 3), 3)
            wms /= max(args.steps // 2, 3)
        whole = {"what": "cer_mvs_b200.raft.RAFT.forward: fp32 images in (0..255), normalisation + fnet x%d + cnet "
                         "(hand-written mma.sync convs, instance norm) + hot path, images resident" % (Vc + 1),
                 "ms_per_depth_map": wms, "depth_maps_per_s": 1e3 / wms,
                 "encoder_gflop_per_depth_map": 71.0 * (Vc + 2)}
        if not args.no_reference_gpu:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            try:
                import refrun
                if refrun.available("gpu"):
                    ref = refrun.import_reference("gpu")
                    refrun.restore_reference_classes()
                    rmodel = ref.raft.RAFT(cascade=CASCADE, test_mode=True)
                    rmodel.load_state_dict(sdw, strict=True)
                    rmodel = rmodel.to(dev).eval()
                    sc64 = torch.tensor([1.0], dtype=torch.float64, device=dev)
                    with torch.no_grad():
                        rms = timed(lambda: rmodel(d_images.clone(), scene.d_poses.clone(), scene.d_K.clone(), scale=sc64),
                                    2, 1) / 2
                    whole["reference_ms_per_depth_map"] = rms
                    whole["speedup_vs_reference_gpu"] = rms / wms
                    del rmodel
            except Exception as e:  # noqa: BLE001
                whole["reference_unavailable"] = repr(e)[:200]
        del model, d_images
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and cfg == "cfg2":
        # one bounded sample of ~10-20 s of CPU work: 24 of 296 rows, 1+1 iterations
        cores = host_threads()
        full, desc, kind = cpu_sample(rows=3 * args.ref_rows, iters=(1, 1))
        cpu = {"value": 1.0 / full, "unit": "depth-maps/s", "cores": cores, "kind": kind, "sample": desc}

    # ---- the reference's GPU path on this same B200: its Python, its kernel, real autocast ----
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_reference_gpu and cfg == "cfg2":
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        try:
            import bench_ref
            import refrun
            if refrun.available("gpu"):
                r = bench_ref.reference_gpu(Hc, Wc, Vc, scene.sc, sd, {"16+16": CASCADE, "8+8": CASCADE_8}, dev)
                ref_gpu = {
                    "what": "UNMODIFIED reference on this GPU: core/raft.py:75-108 with its own CorrBlock / UpdateBlock, "
                            "its own alt_cuda_corr kernel (sm_100 build), torch.cuda.amp.autocast; stub encoders; "
                            "inputs resident; CUDA events",
                    "ms_per_depth_map_16_16": r["16+16"], "depth_maps_per_s_16_16": 1e3 / r["16+16"],
                    "ms_per_depth_map_8_8": r["8+8"], "depth_maps_per_s_8_8": 1e3 / r["8+8"],
                    "speedup_16_16": r["16+16"] / (ms / args.steps),
                    "speedup_8_8": r["8+8"] / ms8 if ms8 else None, "north_star_target": ">= 4x",
                    "corr_kernels": bench_ref.corr_kernel_legs(Hc, Wc, Vc, scene.sc, dev),
                }
            else:
                ref_gpu = {"unavailable": "baseline/_ref or oracle/_ref did not travel"}
        except Exception as e:  # noqa: BLE001   the checker must never break the product's line
            ref_gpu = {"unavailable": repr(e)[:300]}

    if rank == 0:
        if sharded_headline:
            par, scaling = f"one image sharded over {world} GPUs ((view, hypothesis) units, 1 NCCL all-reduce per stage)", "strong"
        else:
            par, scaling = (f"replicas x{world}" if world > 1 else "single GPU"), "weak"
        line = {
            "metric": METRIC if cfg == "cfg2" else f"depth-maps/sec ({CONFIGS[cfg][4]}, 32 iters)",
            "value": value, "unit": "depth-maps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "parallelism": par,
                       "l2": "inputs larger than L2: fp16 feature maps (%d MB) + cost volume + update workspace are "
                             "re-streamed every step (L2 = 126 MB)" % ((Vc + 1) * px * 128 // 1000000),
                       "engine": "cer_plan, CUDA graph per cascade stage",
                       "scope": SCOPE},
            "e2e": {"value": e2e, "unit": "depth-maps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "sync_call_ms": sync_ms, "api": e2e_api},
            "gpu_launches": launches,
            "clocks": clk, "roofline": roof, "build_roofline": build_roof, "lookup_roofline": lookup_roof,
            "cpu_baseline": cpu, "reference_gpu": ref_gpu, "whole_forward": whole, "kernels": kernels,
        }
        if ms8:
            line["iters_8_8"] = {"value": world * 1e3 / ms8, "unit": "depth-maps/s", "ms_per_step": ms8,
                                 "note": "same workload with the reference's default 8+8 iterations"}
        if sharded:
            line["sharded" if not sharded_headline else "replicas"] = sharded
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-rows", type=int, default=8, help="rows of the 296-row grid in one CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference-on-this-GPU leg")
    ap.add_argument("--no-whole-forward", action="store_true", help="skip the whole-RAFT.forward leg (encoders + hot path)")
    ap.add_argument("--profile-step", action="store_true", help="ncu helper: warm-up + one eager step only")
    ap.add_argument("--conv-variant", type=int, default=None, help="cer_set_conv_variant (A/B experiments)")
    ap.add_argument("--build-variant", type=int, default=None, help="cer_set_build_variant (A/B experiments)")
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configuration")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
