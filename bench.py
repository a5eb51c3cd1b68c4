#!/usr/bin/env python
"""bench.py -- frame-pairs/sec of the FlowTrack per-frame-pair hot path (pose + flow + box propagation) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU path on the host cores

N > 1 is launched by torchrun (one rank per GPU, NCCL); every rank runs the same per-GPU workload on its own shard of
frame pairs (weak scaling) and the only collective is the all-gather of the per-person key-point records.

Workload (BASELINE.json config 4, P = 4 people per frame as fixed by SURVEY.md 8d): per GPU and step, 16 frame pairs
= FlowNet2S on [16,3,2,384,512] + PoseResNet-50 on the 64 person crops [64,3,256,192] + heat-map decode + flow
look-up box propagation, synthetic inputs, random-init weights.  A "step" is one such batch.

Keys beyond the base contract: `roofline` (dominant kernel = the tcgen05 implicit-GEMM convolution: algorithmic
FLOPs of all its launches in a step / their summed CUDA-event durations, against the measured dense bf16 GEMM
peak), `kernels` (the same for the other kernels on the path incl. the FlowNetC correlation at BASELINE config 3),
`cpu_baseline` (the oracle port of the reference path on the host cores, bounded sample), `e2e` (host buffers in,
host results out, through FramePairPipeline.submit()/collect(), two batches in flight).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame-pairs/sec (pose+flow+warp) at 256x192/384x512"
UNIT = "frame-pairs/s"
PAIRS, PEOPLE = 16, 4
FRAME_HW, CROP_HW = (384, 512), (256, 192)
GFLOP_PER_PAIR = 31.58 + PEOPLE * 10.853          # FlowNet2S pair + 4 x R50 crop


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tensor_burst=float(p["bf16_tflops"]), tensor=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference path (reference arm and cpu_baseline)
# --------------------------------------------------------------------------------------------------------
def cpu_reference_runner(pairs_per_step):
    """Returns (step_fn, cores, description).  step_fn() runs `pairs_per_step` frame pairs of the reference path on the
    host: FlowNet2S forward, box propagation, PoseResNet-50 forward on P crops per pair, final_preds -- the reference's
    own PyTorch CPU kernels through the oracle's restated modules (oracle/{pose,flow,decode}_oracle.py)."""
    import torch
    from flowtrack.pytorch_b200.tracking import synthetic
    from oracle import decode_oracle, flow_oracle, pose_oracle, track_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    pose = pose_oracle.deconv("resnet50", 17, False).eval()
    synthetic.condition_pose_weights(pose, 0)
    flow = flow_oracle.FlowNet2S(None).eval()
    d = synthetic.frame_pair_batch(pairs_per_step, PEOPLE, FRAME_HW, CROP_HW, seed=1)
    boxes_np = d["boxes"].numpy()
    try:
        import cv2  # noqa: F401
        have_cv2 = True
    except Exception:
        have_cv2 = False

    def step():
        with torch.no_grad():
            fl = flow(d["pair"]).numpy()
            boxes = [decode_oracle.box_propagation(d["prev_keypoints"][i * PEOPLE:(i + 1) * PEOPLE].numpy(), fl[i]) for i in range(pairs_per_step)]
            # pose_est's crop loop (net_utils.py:52-57): one warpAffine per person box out of the pair's current frame (BGR)
            frames = d["pair"][:, :, 1].permute(0, 2, 3, 1).numpy()[:, :, :, ::-1]
            crops = np.concatenate([track_oracle.crop_boxes(np.ascontiguousarray(frames[i]), boxes_np[i * PEOPLE:(i + 1) * PEOPLE], CROP_HW, use_cv2=have_cv2)
                                    for i in range(pairs_per_step)])
            hm = pose(torch.from_numpy(crops)).numpy()
            kps, sc = decode_oracle.final_preds(hm, d["centers"].numpy(), d["scales"].numpy(), True)
        return kps, sc, boxes

    return step, cores, "oracle port (torch CPU fp32, %d threads): FlowNet2S 384x512 + %d PoseResNet-50 crops 256x192 + decode + box propagation per pair" % (cores, PEOPLE)


def time_cpu(step, pairs_per_step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return pairs_per_step * steps / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pps = 1
    step, cores, desc = cpu_reference_runner(pps)
    value, ms = time_cpu(step, pps, args.steps, args.warmup)
    sample = "%d frame pair(s) per step (1 FlowNet2S pair + %d R50 crops + decode + box propagation), %d timed steps" % (pps, PEOPLE, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(1, pps), "gpu_launches": 0,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "what": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def workload_config(world, pairs_per_gpu):
    return {"workload": "config4 joint pipeline: per GPU and step %d frame pairs = FlowNet2S [%d,3,2,%d,%d] + PoseResNet-50 on %d crops [%d,3,%d,%d] "
                        "(P=%d people/frame, crops cut from the current frame by the warpAffine rule of pose_est) + heat-map decode + flow box propagation" % (pairs_per_gpu, pairs_per_gpu, FRAME_HW[0], FRAME_HW[1],
                                                                                           pairs_per_gpu * PEOPLE, pairs_per_gpu * PEOPLE, CROP_HW[0], CROP_HW[1], PEOPLE),
            "pairs_per_gpu": pairs_per_gpu, "people_per_frame": PEOPLE, "global_pairs": pairs_per_gpu * world,
            "parallelism": "frame pairs sharded over %d GPU(s), one all-gather of key-point records per step" % world,
            "l2": "no explicit flush: a step streams ~75 MB of inputs and several GB of activations, far beyond the 126 MB L2",
            "weights": "random-init (reference init + conditioned BatchNorm/head, seed 0)"}


# --------------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from flowtrack.pytorch_b200 import _lib
    from flowtrack.pytorch_b200.flownet.model import models as flow_models
    from flowtrack.pytorch_b200.pose import models as pose_models
    from flowtrack.pytorch_b200.tracking import pipeline as pl
    from flowtrack.pytorch_b200.tracking import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                               # creates the communicator now, inside the redirected region
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.lib()                                           # fail loudly if the CUDA library is missing

    torch.manual_seed(0)
    pose = synthetic.condition_pose_weights(pose_models.deconv("resnet50", 17, False), 0).cuda()
    flow = flow_models.FlowNet2S(None).cuda()
    pose.precision = flow.precision = args.precision
    pipe = pl.FramePairPipeline(pose, flow, args.pairs, PEOPLE, FRAME_HW, CROP_HW, dev, crops_from_frames=not args.host_crops)
    data = synthetic.frame_pair_batch(args.pairs, PEOPLE, FRAME_HW, CROP_HW, seed=100 + rank)
    pipe.process(data["pair"], data["crops"], data["centers"], data["scales"], data["prev_keypoints"])   # fills static + pinned buffers, captures

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        pipe.step_device()
        return pl.gather_keypoints(pipe.keypoints), pl.gather_keypoints(pipe.boxes)

    pending = []

    def host_step():
        # pipelined host API: submit this step's batch (pinned host -> device on the copy stream, graph replay, results ->
        # pinned host), then collect the previous step's results -- every step's copies are inside the timed region, the
        # copy-in of step i+1 overlaps the kernels of step i.  The all-gather runs on the device-resident records.
        pending.append(pipe.submit())
        out = (pl.gather_keypoints(pipe.keypoints), pl.gather_keypoints(pipe.boxes))
        if len(pending) > 1:
            pipe.collect(pending.pop(0))
        return out

    def host_drain():
        while pending:
            pipe.collect(pending.pop(0))

    def timed(fn, steps, warmup, drain=None):
        for _ in range(warmup):
            fn()
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()                                      # the last step's results are on the host before the clock stops
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps                          # ms per step, max over ranks

    # clocks: nvidia-smi needs ~0.1-0.3 s before its first sample and the timed region of 20 steps is < 0.1 s long, so the sampler
    # starts first, the same device step runs untimed until samples are flowing (pre-roll, identical load), and it keeps sampling
    # through the timed device steps and the timed e2e steps
    sampler = ClockSampler(local) if rank == 0 else None
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.preroll:      # time-based, so NO collective in here: ranks may run different counts
        for _ in range(10):
            pipe.step_device()
        torch.cuda.synchronize()
    ms = timed(device_step, args.steps, args.warmup)
    pipe.stage(data["pair"], data["crops"], data["centers"], data["scales"], data["prev_keypoints"], slot=1)   # slot 0 was staged by process()
    ms_e2e = timed(host_step, args.steps, max(1, args.warmup // 2), drain=host_drain)
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "%.1f s untimed pre-roll of the same device step + the timed device steps + the timed e2e steps (20 ms sampling)" % args.preroll
    total_pairs = args.pairs * world

    out = None
    if rank == 0:
        peaks = load_peaks()
        # ---- per-kernel timing: un-graphed pass, CUDA-event pair around every launch ---------------------
        kern = {}
        conv_ms, conv_flops, other = 0.0, 0.0, {}
        for plan in (pipe.pose_plan, pipe.flow_plan):
            per = plan.time_calls(reps=3)
            for i, (what, _fn, _a, _k) in enumerate(plan.calls):
                if what == "conv2d":
                    conv_ms += per[i]
                    conv_flops += plan.flops[i]
                else:
                    other[what] = other.get(what, 0.0) + per[i]
        n_conv = sum(1 for plan in (pipe.pose_plan, pipe.flow_plan) for c in plan.calls if c[0] == "conv2d")
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        traffic, ncu_note = None, None
        try:     # DRAM bytes per conv launch from the committed ncu --set full capture of the same plans (not measured in this run)
            with open(os.path.join(ROOT, "profiles", "r01j_conv_traffic.json")) as f:
                tr = json.load(f)
            traffic = tr["dram_bytes_per_launch"]
            ncu_note = {"tensor_pipe_active_pct_time_weighted": tr["tensor_pipe_active_pct_time_weighted"], "source": tr["source"]}
        except Exception:
            pass
        roof = {"kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv/deconv, %d launches per step)" % n_conv, "bound": "tensor",
                "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": ach / peaks["tensor"], "traffic": traffic,
                "traffic_unit": "DRAM bytes (read+write) per launch, averaged over the conv launches of a step", "ncu": ncu_note,
                "algorithmic_flop_per_launch": conv_flops / n_conv, "ms_per_launch": conv_ms / n_conv,
                "algorithmic_gflop_per_step": conv_flops / 1e9, "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / ms,
                "executed_over_algorithmic": 3.0 if args.precision == "split" else 1.0,
                "peak_source": peaks["source"] + ", sustained dense bf16 GEMM"}
        kern["other_kernels_ms_per_step"] = other
        kern["correlation"] = bench_correlation(torch, _lib, peaks)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            pps = 1
            step, cores, desc = cpu_reference_runner(pps)
            v, _ = time_cpu(step, pps, args.cpu_steps, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d timed steps of %d frame pair (1 FlowNet2S pair + %d R50 crops + decode + box propagation) after 1 warm-up" % (args.cpu_steps, pps, PEOPLE),
                   "what": desc}
        out = {
            "metric": METRIC, "value": total_pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32-grade (fp16 hi/lo split operands, 3 tcgen05 MMAs per product, fp32 accumulate)" if args.precision == "split" else "fp16 operands, fp32 accumulate",
            "data": "synthetic", "config": workload_config(world, args.pairs),
            "gflop_per_pair": GFLOP_PER_PAIR, "algorithmic_tflops": total_pairs * GFLOP_PER_PAIR / ms,
            "e2e": {"value": total_pairs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "api": "FramePairPipeline.submit()/collect(): per step pinned host buffers -> device (copy stream), CUDA-graph replay, results -> pinned host; "
                           "two slots in flight so the copy-in of step i+1 overlaps the kernels of step i"},
            "gpu_launches": pipe.launches_per_step * args.steps, "gpu_launches_per_step": pipe.launches_per_step,
            "clocks": clocks, "roofline": roof, "kernels": kern, "cpu_baseline": cpu,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def bench_correlation(torch, _lib, peaks):
    """FlowNetC correlation at BASELINE config 3: two [32,256,48,64] maps -> [32,441,48,64] fp32; algorithmic bytes =
    inputs once + output once at 4 B per element = 11.71 MB per pair.  Headline = the tcgen05 banded-GEMM
    kernel on the conv stack's split-fp16 NHWC activations (hi+lo = 4 B per element, the layout it reads inside FlowNetC);
    also timed: the same kernel behind the reference's fp32 NCHW signature (re-layout into the scratch buffer + kernel)
    and the CUDA-core kernel."""
    from flowtrack.pytorch_b200 import engine
    B, C, H, W = 32, 256, 48, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(B, C, H, W, device="cuda", generator=g)
    b = torch.randn(B, C, H, W, device="cuda", generator=g)
    o = torch.empty(B, 441, H, W, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    L = _lib.lib()
    sp = lambda: torch.cuda.current_stream().cuda_stream
    acts = []
    for t in (a, b):
        act = engine.Act(B, H, W, C, "cuda")
        st = t.stride()
        _lib.check(L.ft_pack_image(t.data_ptr(), st[0], st[1], st[2], st[3], C, 1.0, act.struct(), 0, 0, sp()), "pack_image")
        acts.append(act)
    nws = L.ft_correlation_workspace_bytes(B, C, H, W)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    fns = {
        "tcgen05": lambda: _lib.check(L.ft_correlation_act(acts[0].struct(), acts[1].struct(), o.data_ptr(), 1, _lib.PREC_SPLIT, sp()), "correlation_act"),
        "tcgen05_planar": lambda: _lib.check(L.ft_correlation_act_planar(acts[0].struct(), acts[1].struct(), o.data_ptr(), 1, _lib.PREC_SPLIT, sp()), "correlation_act_planar"),
        "tcgen05_fast": lambda: _lib.check(L.ft_correlation_act(acts[0].struct(), acts[1].struct(), o.data_ptr(), 1, _lib.PREC_FAST, sp()), "correlation_act"),
        "tcgen05_nchw_abi": lambda: _lib.check(L.ft_correlation_forward_ws(a.data_ptr(), b.data_ptr(), o.data_ptr(), B, C, H, W, 20, 1, 20, 1, 2, 1,
                                                                           _lib.PREC_SPLIT, ws.data_ptr(), nws, sp()), "correlation_ws"),
        "simt": lambda: _lib.check(L.ft_correlation_forward(a.data_ptr(), b.data_ptr(), o.data_ptr(), B, C, H, W, 20, 1, 20, 1, 2, 1, sp()), "correlation"),
    }
    res = {}
    for name, fn in fns.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10 if name != "simt" else 3):
            flush.zero_()                                    # evict the inputs from L2 between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = statistics.median(ts)
    ms = res["tcgen05"]
    byts = (2 * B * C * H * W + B * 441 * H * W) * 4
    ach = byts / (ms * 1e-3) / 1e9
    return {"kernel": "corr_pair_kernel (tcgen05 cta_group::2 banded GEMM, split-fp16 NHWC [32,48,64,256] x2 -> fp32 NCHW [32,441,48,64], LeakyReLU fused), L2 flushed between launches",
            "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "ms": ms, "algorithmic_bytes": byts,
            "gflop": 2 * B * 441 * C * H * W / 1e9,
            "ms_parity_planar_output": res["tcgen05_planar"], "gbs_parity_planar_output": byts / (res["tcgen05_planar"] * 1e-3) / 1e9,
            "note_parity_planar": "same kernel writing the pipeline-internal layout [B,441,4,H/2,W/2] (dense 128-byte warp stores) that FlowNetC consumes through ft_pack_corr_planar",
            "ms_fast_precision": res["tcgen05_fast"], "ms_fp32_nchw_abi_with_relayout": res["tcgen05_nchw_abi"],
            "ms_cuda_core_kernel": res["simt"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS, help="frame pairs per GPU per step")
    ap.add_argument("--precision", default="split", choices=["split", "fast"])
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preroll", type=float, default=0.6, help="seconds of untimed steps while nvidia-smi starts sampling")
    ap.add_argument("--host-crops", action="store_true", help="take ready-made person crops from the host instead of cutting them on the GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
