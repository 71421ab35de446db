#!/usr/bin/env python
"""bench.py -- benchmark of the B200 box-op hot path (contract in the task statement).

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU restatement of the reference path

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[2] -- every box op of one Faster R-CNN R50-FPN training
step (anchors, RPN top-k + decode + NMS 0.7 -> proposals, RPN targets, RCNN targets, ROIAlign 7x7 forward + backward),
16 images -- the one config that contains IoU, matching, NMS and ROIAlign, which is what the metric names.  The `configs`
object of the same JSON line carries configs[0..4] (c1..c5), each with images/s, its dominant kernel's roofline (or
pair-test rate) and its own end-to-end figure.

Scaling: images are sharded across ranks with no collective on the data path.  Default `--scaling strong`:
the config's batch (16 / 16 / 64 / 8 images; config 1 is a single image and runs as one replica per rank) is split over
the N ranks, so per-GPU work shrinks with N; `--scaling weak` gives every rank the full batch.  At N > 1 the strong run
also reports the headline at the weak size (`weak_scaling`).  One "step" = one pass of the pipeline over the rank's
images, replayed as ONE CUDA graph; timing = per-step CUDA events on the launching stream, max over ranks, L2 flushed
between steps (outside the event pairs).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "box-op images/sec (IoU+match+NMS+ROIAlign) @1/2/4/8 B200; % of HBM roofline"
UNIT = "images/s"
FALLBACK_HBM_GBS = 6650.0
HEADLINE = "c3"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def profile_traffic():
    """ncu dram bytes per launch per kernel, committed under profiles/ (NOT measured in this run)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        """windows: list of (t0, t1) perf_counter intervals that were under load."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [s for (t, s) in self.samples if any(a <= t <= b for a, b in windows)] or [s for (_, s) in self.samples]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from basedet_b200 import _lib, benchmarks as BM, distributed as D, ops

    _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    peak, peak_src = measured_peak()
    traffic = profile_traffic()
    warm = max(args.warmup, 3)
    windows = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def shard(name, scaling):
        n = BM.CONFIG_BATCH[name]
        if args.simulate_world > 1 and world == 1:
            # analysis aid: run rank 0's shard of a `simulate_world`-rank strong-scaling job on this one GPU
            lo, hi = D.shard_range(n, 0, args.simulate_world)
            return list(range(lo, max(hi, lo + 1))), max(hi - lo, 1), "strong-shard-of-%d" % args.simulate_world
        if scaling == "weak" or n < world:
            # weak: every rank runs the full batch (its own images); configs smaller than the world: one replica per rank
            return list(range(rank * n, (rank + 1) * n)), n * world, ("weak" if scaling == "weak" else "replicas")
        lo, hi = D.shard_range(n, rank, world)
        return list(range(lo, hi)), n, "strong"

    def measure(name, scaling, steps, profile=True):
        images, total_images, mode = shard(name, scaling)
        arm = BM.ARMS[name](images, dev)
        if hasattr(arm, "finish_setup"):
            arm.finish_setup()
        graphed = arm.capture() if not args.eager else False
        # ---- device-resident timing
        for _ in range(warm):
            arm.step()
        barrier()
        evs = []
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            arm.step()
            e.record()
            evs.append((s, e))
        barrier()
        t1 = time.perf_counter()
        per_step = [s.elapsed_time(e) for s, e in evs]
        dev_ms = float(sum(per_step))
        # ---- end to end: pinned host inputs -> H2D -> step -> result summary -> D2H, every step
        res_h = torch.empty((steps,) + tuple(arm.summary.shape), dtype=arm.summary.dtype).pin_memory()
        for _ in range(3):
            arm.e2e_step(res_h[0])
        barrier()
        te0 = time.perf_counter()
        for i in range(steps):
            arm.e2e_step(res_h[i])
        barrier()
        te1 = time.perf_counter()
        windows.append((t0, te1))
        e2e_ms = (te1 - te0) * 1e3
        # ---- per-kernel durations: eager launches bracketed by the library's event hooks (outside the timed regions)
        kernels, launches_per_step = {}, None
        if profile:
            n_prof = 3
            arm.eager()
            torch.cuda.synchronize()
            ops.profile_begin()
            for _ in range(n_prof):
                flush.zero_()
                arm.eager()
            torch.cuda.synchronize()
            rep = ops.profile_report()
            ops.profile_end()
            launches_per_step = sum(n for _, n in rep.values()) / n_prof
            tot = sum(ms for ms, _ in rep.values()) or 1.0
            for kname, (ms, n) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
                d = {"avg_us": ms / n * 1e3, "launches_per_step": n / n_prof, "share_of_kernel_time": ms / tot}
                if kname in arm.kernel_bytes:
                    gbs = arm.kernel_bytes[kname] / (ms / n * 1e-3) / 1e9
                    d.update(algorithmic_bytes_per_launch=arm.kernel_bytes[kname], GBps=gbs, frac_of_peak=gbs / peak)
                if kname in arm.kernel_units:
                    unit, units = arm.kernel_units[kname]
                    d[unit] = units / (ms / n * 1e-3)
                kernels[kname] = d
        # ---- max over ranks
        times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dev_ms_max, e2e_ms_max = float(times[0]), float(times[1])
        rec = {
            "workload": arm.workload, "images_total": total_images, "images_this_rank": len(images), "scaling": mode,
            "value": total_images * steps / (dev_ms_max * 1e-3), "unit": UNIT, "ms_per_step": dev_ms_max / steps,
            "ms_per_step_min": float(min(per_step)), "cuda_graph": bool(graphed),
            "e2e": {"value": total_images * steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": arm.h2d_bytes, "d2h_bytes_per_step": arm.d2h_bytes,
                    "note": "pinned host inputs -> H2D -> step (CUDA graph replay) -> result summary -> D2H every step; wall "
                            "clock between device syncs; " + arm.resident_note},
            "gpu_launches_per_step": launches_per_step, "kernels": kernels,
        }
        if not graphed and getattr(arm, "capture_error", None):
            rec["cuda_graph_error"] = arm.capture_error
        dom = arm.dominant if arm.dominant in kernels else (next(iter(kernels)) if kernels else None)
        if dom:
            k = kernels[dom]
            rec["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": k.get("GBps"), "peak": peak, "unit": "GB/s",
                               "frac": k.get("frac_of_peak"), "traffic": traffic.get(dom),
                               "traffic_source": "profiles/traffic.json (ncu --set full capture, not measured in this run)"
                               if traffic.get(dom) is not None else None,
                               "algorithmic_bytes_per_launch": k.get("algorithmic_bytes_per_launch"),
                               "avg_launch_ms": k["avg_us"] / 1e3, "peak_source": peak_src}
            for unit in ("pair-evaluations/s",):
                if unit in k:
                    rec["roofline"][unit] = k[unit]
        if name == "c3":
            rec["roi_level_histogram"] = arm.level_hist
            rec["roi_footprint_union_bytes"] = arm.union_bytes
            rec["pyramid_bytes"] = arm.pyramid_bytes
        del arm
        torch.cuda.empty_cache()
        return rec

    sampler = ClockSampler(local) if rank == 0 else None
    names = [n for n in ("c3", "c1", "c2", "c4", "c5") if not args.only or n in args.only.split(",")]
    configs = {}
    for n in names:
        configs[n] = measure(n, args.scaling, args.steps)
    weak = None
    if world > 1 and args.scaling == "strong" and HEADLINE in configs:
        w = measure(HEADLINE, "weak", args.steps, profile=False)
        weak = {"value": w["value"], "unit": UNIT, "ms_per_step": w["ms_per_step"], "images_total": w["images_total"],
                "e2e": w["e2e"]["value"]}
    # keep a load running until the clock sampler has seen >= 0.5 s of it
    if sampler is not None and windows and sum(b - a for a, b in windows) < 0.6:
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.6:
            flush.zero_()
        torch.cuda.synchronize()
        windows.append((t0, time.perf_counter()))
    clocks = sampler.stop(windows) if sampler else None

    if rank == 0:
        head = configs.get(HEADLINE) or next(iter(configs.values()))
        out = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": head["scaling"] if head["scaling"] != "replicas" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": head["workload"], "images_total": head["images_total"],
                       "images_per_gpu": head["images_this_rank"],
                       "sharding": "images across ranks (distributed.shard_range), no collective on the data path",
                       "l2": "256 MB memset between timed steps (outside the per-step CUDA event pairs)",
                       "timing": "sum of per-step CUDA event intervals on the launching stream, max over ranks; the step is "
                                 "one CUDA graph replay",
                       "other_configs": "configs.c1 .. configs.c5 of this line"},
            "e2e": head["e2e"],
            "gpu_launches": int(round((head["gpu_launches_per_step"] or 0) * args.steps)),
            "roofline": head.get("roofline"),
            "clocks": clocks,
            "configs": configs,
        }
        if weak is not None:
            out["weak_scaling"] = weak
        out["cpu_baseline"] = cpu_baseline(HEADLINE if HEADLINE in configs else names[0])
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- CPU arm
class CpuArm:
    """The reference path on the host cores.  BaseDet is pure Python over MegEngine, which cannot be installed offline,
    so the timed code is oracle/cpu_arms.py: the numpy + C (oracle/c/oracle.c, pthread-parallel, materialised matrices,
    one pass per reference op) restatement that tests/ pin to the reference's own source.  One call = ONE image."""

    def __init__(self, name):
        from basedet_b200 import workloads as W
        from oracle import c_oracle, cpu_arms, ref_ops as R
        self.name, self.W, self.CA, self.R = name, W, cpu_arms, R
        self.cores = c_oracle.num_threads()
        rng = np.random.default_rng(0)
        if name == "c3":
            fs = [(-(-W.FRCNN_HW[0] // s), -(-W.FRCNN_HW[1] // s)) for s in W.FRCNN_RCNN_STRIDES]
            # one image's pyramid / dout (activation values do not steer control flow: reused for every sampled image)
            self.feats = [rng.standard_normal((1, W.FRCNN_CHANNELS, h, w), dtype=np.float32) for h, w in fs]
            self.dout = rng.standard_normal((W.FRCNN_NUM_ROIS, W.FRCNN_CHANNELS, 7, 7), dtype=np.float32)
            self.inputs = lambda i: W.frcnn_image(i % 16)
            self.what = "anchors + RPN proposals + RPN targets + RCNN targets + ROIAlign fwd + bwd (fp64-accumulated)"
        elif name == "c2":
            sizes = W.retinanet_level_sizes(800, 800)
            self.sizes = sizes
            self.scratch = c_oracle.TargetScratch(100, sum(h * w * 9 for h, w in sizes))
            self.inputs = lambda i: W.target_assign_batch(1, 100, 800, 800, seed0=100 + i % 16)[0][0]
            self.what = "anchors + materialised IoU + Matcher + class labels + BoxCoder.encode"
        elif name == "c1":
            self.sizes = W.retinanet_level_sizes(800, 800)
            self.inputs = lambda i: W.retinanet_image(0)
            self.what = "anchors + sigmoid filter + top-k + decode + batched NMS + scale / clip"
        elif name == "c4":
            self.sizes = W.retinanet_level_sizes(*W.FCOS_HW)
            self.inputs = lambda i: W.fcos_image(i % 64)
            self.what = "points + FCOS score filter + top-k + PointCoder.decode + batched NMS + scale / clip"
        elif name == "c5":
            self.inputs = lambda i: W.stress_image(i % 8)
            self.what = "500 x 200k IoU + Matcher + single-class NMS over 100k boxes"
        else:
            raise ValueError(name)

    def image(self, inp):
        W, CA, R = self.W, self.CA, self.R
        if self.name == "c3":
            return CA.frcnn_image_chain(inp, self.feats, self.dout)
        if self.name == "c2":
            anchors = np.concatenate(R.default_anchors(self.sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS,
                                                       W.RETINANET_STRIDES, W.RETINANET_OFFSET))
            return CA.retinanet_targets_image(anchors, inp, self.scratch)
        if self.name == "c1":
            anc = R.default_anchors(self.sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, W.RETINANET_OFFSET)
            return CA.dense_image_postprocess(inp["logits"], inp["offsets"], anc, inp["im_info"])
        if self.name == "c4":
            pts = R.anchor_points(self.sizes, 1, W.RETINANET_STRIDES, 0.5)
            return CA.dense_image_postprocess(inp["logits"], inp["offsets"], pts, inp["im_info"], 0.05, W.FCOS_NMS_THR, 100,
                                              1000, ctrness_l=inp["ctrness"])
        return CA.stress_image_ops(inp)


def cpu_baseline(name, budget_s=15.0, max_images=64):
    arm = CpuArm(name)
    inputs = [arm.inputs(i) for i in range(4)]
    arm.image(inputs[0])
    n_img, t0 = 0, time.perf_counter()
    while True:
        arm.image(inputs[n_img % len(inputs)])
        n_img += 1
        if time.perf_counter() - t0 > budget_s or n_img >= max_images:
            break
    dt = time.perf_counter() - t0
    return {"value": n_img / dt, "unit": UNIT, "cores": arm.cores, "kind": "port", "config": name,
            "sample": "%d images of the same workload, one at a time (%s); numpy + C restatement of the reference op sequence, "
                      "C passes on %d pthreads (MegEngine itself is not installable offline)" % (n_img, arm.what, arm.cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = (args.only or HEADLINE).split(",")[0]
    arm = CpuArm(name)
    inputs = [arm.inputs(i) for i in range(min(args.steps, 16))]
    for i in range(max(args.warmup, 1)):
        arm.image(inputs[0])
    t0 = time.perf_counter()
    for i in range(args.steps):
        arm.image(inputs[i % len(inputs)])  # bounded sample: ONE image of the config's batch per step
    dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = ("1 image per step (%s); numpy + C restatement (oracle/cpu_arms.py, oracle/c/oracle.c) of the reference's "
              "MegEngine op sequence, C passes on %d pthreads; the reference is pure Python over MegEngine, not installable "
              "offline; timed region %.1f s" % (arm.what, arm.cores, dt))
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[%d] (%s), each step = 1 image (bounded sample of the batch)"
                               % (int(name[1]) - 1, name), "path": "cpu-oracle"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


_RESULT_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; see main()."""
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (line + "\n").encode())


def main():
    # Libraries (NCCL prints its version banner, torch warnings ...) write to fd 1 at will; the contract is one JSON
    # line on stdout, so everything else is routed to stderr and the result is written to the saved descriptor.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--only", default="", help="comma-separated subset of c1,c2,c3,c4,c5")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of the CUDA graph replay")
    ap.add_argument("--simulate-world", type=int, default=1,
                    help="analysis aid (1 GPU): time rank 0's shard of an N-rank strong-scaling run")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and args.gpus != world:
        if world == 1 and args.gpus > 1:
            # honour --gpus when started without torchrun: re-launch as one process per GPU
            import socket
            s = socket.socket()
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
            s.close()
            os.dup2(_RESULT_FD, 1)
            os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                       "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                       "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
