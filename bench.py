#!/usr/bin/env python
"""Benchmark of the early-exit segmentation hot path (BASELINE.json metric: early-exit images/sec at
513x513).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl eeseg|reference]

Workload (config.workload): BranchyDeepLabV3 ResNet-50, 3 exits (heads on 1024/2048/2048 channels),
entropy-threshold early-exit inference in bf16 on synthetic VOC-shaped 513x513 images, 4 images per
GPU (global batch 32 on 8 GPUs -> weak scaling). A step = one batch through
EarlyExitEngine.evaluate: backbone sections, tcgen05 exit heads, fused up-sample/softmax/entropy/
argmax gate, per-image exit decision, confusion histogram of the exit taken.

One JSON line on stdout (rank 0): value = images/s with inputs resident in HBM; e2e = the same step
from pinned HOST buffers with H2D of images+targets and D2H of the per-image results inside the
timed region; roofline = the implicit-GEMM conv kernel (dominant) against the measured bf16 peak;
cpu_baseline = the oracle port of the reference's CPU path on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SECTIONS = [16, 3, 1]      # SURVEY.md §3.1 probe: n=2 -> heads on Cin = 1024, 2048
N_CLASSES = 21
IMG = 513                  # int (square) or (H, W)
PER_GPU_BATCH = 4
TAU = 0.5
METRIC = "early_exit_images_per_sec_513"
WORKLOAD_NAME = "synthetic VOC 21-class 513x513"

# --workload: the default is the configuration BASELINE.json's metric is quoted on (configs[1] at one
# GPU's share); "cityscapes" is configs[4]'s shape (full-resolution 1024x2048, 19 classes)
WORKLOADS = {
    "voc513": dict(N_CLASSES=21, IMG=513, PER_GPU_BATCH=4, METRIC="early_exit_images_per_sec_513",
                   WORKLOAD_NAME="synthetic VOC 21-class 513x513"),
    "cityscapes": dict(N_CLASSES=19, IMG=(1024, 2048), PER_GPU_BATCH=2, METRIC="early_exit_images_per_sec_1024x2048",
                       WORKLOAD_NAME="synthetic Cityscapes-shaped 19-class 1024x2048"),
}


def set_workload(name):
    globals().update(WORKLOADS[name])


def img_hw(img=None):
    img = IMG if img is None else img
    return (img, img) if isinstance(img, int) else tuple(img)


def synth_batch(rank, n, img=None, n_classes=None):
    """seed 1234+rank, randn images, blocky labels at 1/16 resolution, 5 % void."""
    import torch
    import torch.nn.functional as F
    H, W = img_hw(img)
    n_classes = N_CLASSES if n_classes is None else n_classes
    g = torch.Generator().manual_seed(1234 + rank)
    X = torch.randn(n, 3, H, W, generator=g)
    low = torch.randint(0, n_classes, (n, 1, (H + 15) // 16, (W + 15) // 16), generator=g)
    y = F.interpolate(low.float(), size=(H, W), mode="nearest").long()
    void = torch.rand(n, 1, H, W, generator=g) < 0.05
    y = torch.where(void, torch.full_like(y, n_classes), y)
    return X, y


def head_flops(h, w, n, cins, n_classes=None):
    """Nominal dense FLOPs (2*M*N*K, no discount for taps in the zero padding) of the implicit-GEMM
    launches of one step: per head 1x1 + 3 atrous 3x3 (Cin->256), projection (4*256->256, the pooled
    branch enters as a shift), 3x3 256->256, final 1x1 256->Cpad."""
    n_classes = N_CLASSES if n_classes is None else n_classes
    M = n * h * w
    cp = (n_classes + 15) // 16 * 16
    tot = 0
    for cin in cins:
        tot += 2 * M * 256 * cin * (1 + 3 * 9)
        tot += 2 * M * 256 * (4 * 256)
        tot += 2 * M * 256 * (9 * 256)
        tot += 2 * M * cp * 256
    return tot


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="eeseg_clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv launch from the committed ncu capture of one
    step of the default workload (profiles/r02_conv_traffic.json); None for other workloads."""
    if METRIC != "early_exit_images_per_sec_513":
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_conv_traffic.json")) as f:
            return json.load(f)["traffic_bytes_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_images=None):
    """The reference's CPU path on all host threads, `sample_images` images per step (default: the config's per-GPU
    batch). When the unmodified reference is staged (oracle/_ref, see oracle/build_ref.py; /root/reference in the build
    container) and the workload is the 21-class one its constructor hard-codes: the reference's OWN
    `from_deepv3_new.branchyDeepv3(base, 'deeplabv3_resnet50', 2, 513, count_branches=False)` (its FLOP-quantile
    placement then gives the sections 16/3/1 of config.workload) and `eval_br_ent.br_evaluator` over a batch-1 loader
    (the reference needs batch 1), kind = "reference". Otherwise the oracle port (oracle/model_port), kind = "port"."""
    import torch
    from oracle import ref_import
    torch.set_num_threads(os.cpu_count() or 1)
    n = PER_GPU_BATCH if sample_images is None else sample_images
    X, y = synth_batch(0, n)
    H, W = img_hw()
    if ref_import.available() and N_CLASSES == 21 and H == W:
        import warnings
        import torchvision
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fd_new, ebe = ref_import.load("from_deepv3_new", "eval_br_ent")
        base_path = os.path.join(tempfile.mkdtemp(prefix="eeseg_ref_"), "base_r50.pth")
        torch.manual_seed(0)
        torch.save(torchvision.models.segmentation.deeplabv3_resnet50(
            weights=None, weights_backbone=None, num_classes=21, aux_loss=True), base_path)
        net = fd_new.branchyDeepv3(base_path, "deeplabv3_resnet50", 2, H, count_branches=False).eval()
        assert [len(sec) for sec in net.base_model] == SECTIONS, [len(sec) for sec in net.base_model]
        loader = [(X[k:k + 1], y[k:k + 1]) for k in range(n)]
        cpu = torch.device("cpu")

        def step():
            return ebe.br_evaluator(net, 3, N_CLASSES, loader, cpu, TAU)
        kind = "reference"
        how = ("UNMODIFIED reference (" + ("oracle/_ref" if ref_import.STAGED else ref_import.REFERENCE_DIR) +
               "): from_deepv3_new.branchyDeepv3 + eval_br_ent.br_evaluator, batch-1 loader")
    else:
        from oracle import model_port
        net = model_port.build_port(SECTIONS, seed=0, num_classes=N_CLASSES).eval()

        def step():
            return model_port.evaluate_batch_cpu(net, X, y, N_CLASSES, TAU)
        kind = "port"
        how = "oracle/model_port.evaluate_batch_cpu (the reference is not staged or hard-codes 21 classes)"
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"value": n / mean, "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": kind, "ms_per_step": mean * 1e3,
            "sample": f"{n} images of {H}x{W} per step x {steps} steps (+{warmup} warm-up), fp32 CPU, 3 exits, {how}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # the config's own batch per step; K and W as asked, bounded so the run ends within a few minutes
    # (a step of 4 batch-1 images through the unmodified reference is ~5-10 s on 8-16 cores)
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    cb = cpu_reference_run(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus):
    return {"workload": "BranchyDeepLabV3 ResNet-50, 3 exits (sections 16/3/1, heads Cin 1024/2048/2048), "
                        f"entropy-threshold early-exit inference, {WORKLOAD_NAME}",
            "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus, "tau": TAU,
            "parallelism": f"dp{n_gpus}", "l2": "flushed between timed steps (256 MiB write)"}


def kernel_intervals(fn, reps, flush):
    """GPU-side start/end of every kernel of `reps` calls of fn() (CUDA-graph replays included), from CUPTI's
    activity records (torch.profiler). Returns a list (one per call) of [(name, start_us, end_us), ...]; the L2-flush
    fill that precedes each call is the separator and is dropped."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            flush.fill_(1)
            fn()
        torch.cuda.synchronize()
    path = tempfile.mktemp(prefix="eeseg_trace_", suffix=".json")
    prof.export_chrome_trace(path)
    with open(path) as f:
        tr = json.load(f)
    os.unlink(path)
    ev = sorted(((e["ts"], e["ts"] + e["dur"], e["name"]) for e in tr["traceEvents"]
                 if e.get("cat") == "kernel" and "dur" in e), key=lambda t: t[0])
    calls, cur = [], None
    for a, b, name in ev:
        if "FillFunctor" in name and (b - a) > 20:          # the 256 MiB flush (tiny fills inside a step are ~2 us)
            cur = []
            calls.append(cur)
        elif cur is not None:
            cur.append((name, a, b))
    return [c for c in calls if c]


def busy_us(intervals):
    """Length of the union of [start, end) intervals."""
    tot, end = 0.0, None
    for a, b in sorted(intervals):
        if end is None or a > end:
            tot += b - a
            end = b
        elif b > end:
            tot += b - end
            end = b
    return tot


def dist_max(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def leg_cityscapes_sweep(dev, rank, world, steps, warmup, barrier):
    """BASELINE configs[4]: full-resolution Cityscapes-shaped evaluation (19 classes, 1024x2048, 2 images per GPU per
    step), ONE forward per batch resolving a sweep of exit thresholds (engine.ThresholdSweep — what eval_br_ent.py:38-84
    + CLI -t :96 needs one whole run per tau for), with the integer confusion-matrix all-reduce INSIDE the timed region."""
    import torch
    from ee_semantic_segmentation_b200.engine import ThresholdSweep
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    C, hw, B = 19, (1024, 2048), 2
    torch.manual_seed(0)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=SECTIONS, pretrained=False, num_classes=C).to(dev).eval()
    net.strict_kernels = True
    Xh, yh = synth_batch(rank, B, img=hw, n_classes=C)
    Xd, yd = Xh.to(dev), yh.to(dev)
    # what the e2e leg uploads: bf16 images (the stem rounds to bf16 first: identical results) and uint8 labels
    Xh, yh = Xh.to(torch.bfloat16).pin_memory(), yh.to(torch.uint8).pin_memory()
    probe = ThresholdSweep(net, C, [0.5])
    sc = probe.update(Xd, yd)[:2].float().cpu()
    taus = [round(0.1 * k, 1) for k in range(1, 10)] + [float(sc[0].median()), float(sc[1].median())]
    sweep = ThresholdSweep(net, C, taus)
    for _ in range(max(warmup, 2)):
        sweep.update(Xd, yd)
    out = {}
    for name, inputs in (("value", lambda: (Xd, yd)),
                         ("e2e", lambda: (Xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)))):
        for _ in range(3):                       # per input dtype: the forward graph is captured on second sight
            sweep.update(*inputs())
        sweep.cm.zero_(); sweep.counts.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(steps):
            sweep.update(*inputs())
        sweep.all_reduce()                     # int64 [T, E+1, C+1, C] + counters, NCCL, once per sweep
        b.record()
        barrier()
        (ms,) = dist_max([a.elapsed_time(b)], dev, world)
        out[name] = (B * world * steps / (ms * 1e-3), ms / steps)
    res = sweep.results()
    assert all(r["out_gl"] == B * world * steps for r in res)
    del sweep, probe, net
    torch.cuda.empty_cache()
    return {"metric": "tau_sweep_images_per_sec_1024x2048", "value": out["value"][0], "unit": "images/s",
            "ms_per_step": out["value"][1], "steps": steps, "per_gpu_batch": B, "n_taus": len(taus), "n_classes": C,
            "e2e": {"value": out["e2e"][0], "unit": "images/s", "ms_per_step": out["e2e"][1],
                    "h2d_bytes_per_step": Xh.numel() * 2 + yh.numel(), "mode": "sequential upload (bf16 images, uint8 labels) + update"},
            "collective": "one int64 all-reduce(SUM) of cm[T,E+1,C+1,C] + counts inside the timed region (per sweep, not per batch)",
            "exits_at_median_tau": {k: res[-2][k] for k in ("b1_count", "b2_count", "count_out")},
            "l2": "working set of a step (50 MB of inputs, GBs of activations) exceeds the 126 MB L2; no flush"}


def leg_train(kind, dev, rank, world, steps, warmup, barrier):
    """BASELINE configs[2] / [3]: one training step (train_funcs.py:12-33) = forward, multi-exit loss, backward, SGD on
    synthetic crops, replayed as one CUDA graph (train_funcs.GraphedTrainStep); data parallel: ONE flat gradient
    all-reduce(AVG) over NCCL captured inside the step graph, i.e. inside the timed region."""
    import torch
    from ee_semantic_segmentation_b200.branchy_seg_losses import LovaszSoftmax
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    from ee_semantic_segmentation_b200.train_funcs import GraphedTrainStep, make_optimizer
    C, img, B = (21, 513, 4) if kind == "ce" else (19, 768, 2)
    torch.manual_seed(0)                                   # identical initial weights on every rank
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, img, sections=SECTIONS, pretrained=False, num_classes=C).to(dev).train()
    net.strict_kernels = True
    opt = make_optimizer(net, lr=1e-3, base_lr=1e-4, buckets=4 if world > 1 else 1)   # eeseg SGD; DP: 4 overlapped buckets
    loss = (BrXEntropyLoss(ignore_index=C, b_reduction="sum", n_exits=3) if kind == "ce"      # main_bradeepv3_ce.py:121
            else LovaszSoftmax(classes="present", ignore=C, n_branches=2))                      # main_bradeepv3.py:121
    Xh, yh = synth_batch(rank, B, img=img, n_classes=C)
    Xh, yh = Xh.pin_memory(), yh.pin_memory()
    Xd, yd = Xh.to(dev), yh.to(dev)
    gstep = GraphedTrainStep(net, loss, opt, Xd, yd)
    for _ in range(max(warmup, 3)):
        l = gstep(Xd, yd)
    out = {}
    for name in ("value", "e2e"):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            if name == "value":
                l = gstep(Xd, yd)
            else:
                lv = float(gstep(Xh, yh))                  # H2D of the batch + D2H of the loss every step
        b.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        (ms,) = dist_max([max(a.elapsed_time(b), wall - 0.5) if name == "e2e" else a.elapsed_time(b)], dev, world)
        out[name] = (B * world * steps / (ms * 1e-3), ms / steps)
    lv = float(l)
    nparam = sum(p.numel() for p in net.parameters())
    gstep.release()
    del gstep, opt, net
    torch.cuda.empty_cache()
    return {"metric": f"train_{kind}_images_per_sec_{img}", "value": out["value"][0], "unit": "images/s",
            "ms_per_step": out["value"][1], "steps": steps, "per_gpu_batch": B, "n_classes": C, "loss": lv,
            "loss_fn": "BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=3)" if kind == "ce"
                       else "BSL.LovaszSoftmax(classes='present', ignore=19, n_branches=2)",
            "e2e": {"value": out["e2e"][0], "unit": "images/s", "ms_per_step": out["e2e"][1],
                    "h2d_bytes_per_step": Xh.numel() * 4 + yh.numel() * 8, "d2h_bytes_per_step": 4},
            "collective": (f"{nparam * 4 / 1e6:.0f} MB of fp32 gradients per step in 4 NCCL all-reduce(AVG) buckets issued from gradient "
                           "hooks on a side stream while the backward runs, all captured in the step graph (inside the timed "
                           "region)") if world > 1 else "none (1 rank)",
            "dtype": "bf16 activations, fp32 master weights / gradients", "optimizer": "eeseg single-launch SGD (momentum 0.9, wd 5e-4, 3 param groups)",
            "l2": "a step touches > 10 GB; no flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="eeseg", choices=["eeseg", "reference"])
    ap.add_argument("--workload", default="voc513", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch of the conv kernel")
    ap.add_argument("--no-overlap-heads", action="store_true", help="early-exit heads in stream order (A/B of the side-stream overlap)")
    ap.add_argument("--legs", default="all", help="comma list of extra workloads after the headline: cityscapes_sweep, "
                                                  "train_ce, train_lovasz; 'all' (default) or 'none'")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args)
    if args.no_pdl:
        os.environ["EESEG_CONV_PDL"] = "0"          # read once when the library is first used

    import torch
    import torch.distributed as dist
    from ee_semantic_segmentation_b200 import _lib, head_plan
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the eeseg kernels have no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    steps = args.steps
    legs = ([] if args.legs == "none" else ["cityscapes_sweep", "train_ce", "train_lovasz"] if args.legs == "all"
            else [l for l in args.legs.split(",") if l])
    if args.workload != "voc513":
        legs = []

    torch.manual_seed(0)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, img_hw()[0], sections=SECTIONS, pretrained=False,
                        num_classes=N_CLASSES).to(dev).eval()
    net.strict_kernels = True                       # a module without an eeseg kernel plan is an error, not a cuDNN call
    eng = EarlyExitEngine(net, N_CLASSES, TAU, use_graph=not args.no_graph)
    eng.overlap_heads = not args.no_overlap_heads
    Xh, yh = synth_batch(rank, PER_GPU_BATCH)
    Xh, yh = Xh.pin_memory(), yh.pin_memory()
    Xd, yd = Xh.to(dev), yh.to(dev)
    # what a bf16 inference stream uploads: bf16 images (the stem rounds fp32 images to bf16 as its first step, so the
    # results are bit-identical) and uint8 labels (void = any value >= C), prepared once like a loader worker would
    Xh16, yh8 = Xh.to(torch.bfloat16).pin_memory(), yh.to(torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, reduce_eng=None):
        """k steps, each bracketed by CUDA events on the launching (current) stream, L2 flushed
        between steps; returns (sum of step ms, wall ms incl. flushes)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        t0 = time.perf_counter()
        torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" profiles these launches only
        for a, b in ev:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.nvtx.range_pop()
        if world > 1 and reduce_eng is not None:
            reduce_eng.all_reduce()           # the sweep's one integer collective (once, not per step)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return sum(a.elapsed_time(b) for a, b in ev), wall

    def step_device():
        eng.evaluate(Xd, yd)

    # eeseg kernel launches and conv FLOPs of one step (counted on an eager step; a graph replay re-issues the same)
    eng_count = EarlyExitEngine(net, N_CLASSES, TAU, use_graph=False)
    eng_count.evaluate(Xd, yd)
    flop_log = []
    head_plan.FLOP_LOG = flop_log
    c0 = _lib.launch_count()
    eng_count.evaluate(Xd, yd)
    launches_per_step = _lib.launch_count() - c0
    head_plan.FLOP_LOG = None
    del eng_count
    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    eng.reset()                                   # exit statistics of the timed steps only
    dev_ms, wall_ms = timed(step_device, steps, eng)
    launches = _lib.launch_count() - l0
    exit_res = eng.results()                      # after the sweep's all-reduce: all ranks' images

    # ---- end to end through the public streaming API: pinned host batches in, per-image results out; uploads, graph
    # replays and read-backs overlap (double-buffered inputs), all inside the timing -------------------------------
    def host_batches(n, X, y):
        for _ in range(n):
            yield X, y

    def run_pipelined(engine, X, y):
        for _ in engine.evaluate_pipelined(host_batches(3, X, y)):
            pass
        barrier()
        t0 = time.perf_counter()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n_out = 0
        for ex, sc in engine.evaluate_pipelined(host_batches(steps, X, y)):
            n_out += ex.numel()
        b2.record()
        if world > 1:
            engine.all_reduce()
        barrier()
        assert n_out == PER_GPU_BATCH * steps
        ms = max(a.elapsed_time(b2), 0.0)
        wall = (time.perf_counter() - t0) * 1e3
        return max(ms, wall - 0.5)     # events bracket the stream work; the wall clock includes the last read-back

    def h2d_alone(engine, X, y):
        Xs, ys = engine.static_inputs(X.shape)
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            Xs.copy_(X, non_blocking=True)
            ys.copy_(y.view_as(ys), non_blocking=True)
        b2.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b2) / 5

    e2e32_ms = h2d32_ms = h2d_ms = None
    if eng.use_graph:
        e2e32_ms = run_pipelined(eng, Xh, yh)                     # the reference loader's dtypes: fp32 images, int64 labels
        h2d32_ms = h2d_alone(eng, Xh, yh)
        eng16 = EarlyExitEngine(net, N_CLASSES, TAU, use_graph=True, input_dtype=torch.bfloat16, target_dtype=torch.uint8)
        e2e_ms = run_pipelined(eng16, Xh16, yh8)
        h2d_ms = h2d_alone(eng16, Xh16, yh8)
        # same decisions and confusion matrices from either input format
        eng.reset(); eng16.reset()
        ra, rb = eng.evaluate(Xd, yd), eng16.evaluate(Xh16.to(dev), yh8.to(dev))
        assert torch.equal(ra["exit"], rb["exit"]) and torch.equal(ra["pred"], rb["pred"]) and torch.equal(eng.cm, eng16.cm)
        del eng16
    else:
        def step_e2e():
            out = eng.evaluate(Xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True))
            return out["exit"].cpu(), out["scores"].cpu()
        for _ in range(2):
            step_e2e()
        e2e_ms, _ = timed(step_e2e, steps)

    # ---- an early-exit operating point: tau between the middle first-exit scores of the batch, so
    # half of the images leave at exit 1; timed through the compute-skipping engine (still-active
    # images are compacted after every gate: later sections run on fewer images)
    sc0 = eng.evaluate(Xd, yd)["scores"][0].float().cpu().sort().values
    mid = (len(sc0) - 1) // 2
    tau_mid = float((sc0[mid] + sc0[mid + 1]) / 2) if len(sc0) > 1 else float(sc0[0]) + 1.0
    eng_skip_eager = EarlyExitEngine(net, N_CLASSES, tau_mid, skip_compute=True)
    for _ in range(3):
        eng_skip_eager.evaluate(Xd, yd)
    skip_eager_ms, _ = timed(lambda: eng_skip_eager.evaluate(Xd, yd), steps)
    # the same engine with one CUDA graph per (exit stage, still-active image count); the host reads the
    # 4-byte active count after each gate and replays the next stage's graph of that size
    eng_skip = EarlyExitEngine(net, N_CLASSES, tau_mid, skip_compute=True, use_graph=not args.no_graph)
    for _ in range(3):
        eng_skip.evaluate(Xd, yd)
    eng_skip.reset()
    skip_ms, _ = timed(lambda: eng_skip.evaluate(Xd, yd), steps)
    skip_counts = [int(v) for v in eng_skip.counts.cpu()]
    skip_px = [int(v) for v in eng_skip.exited_px.cpu()]
    skip_e2e_ms = None
    if eng_skip.use_graph:
        eng_skip16 = EarlyExitEngine(net, N_CLASSES, tau_mid, skip_compute=True, use_graph=True,
                                     input_dtype=torch.bfloat16, target_dtype=torch.uint8)
        for _ in eng_skip16.evaluate_pipelined(host_batches(3, Xh16, yh8)):
            pass
        barrier()
        t0 = time.perf_counter()
        for ex, sc in eng_skip16.evaluate_pipelined(host_batches(steps, Xh16, yh8)):
            pass
        torch.cuda.synchronize()
        skip_e2e_ms = (time.perf_counter() - t0) * 1e3
        del eng_skip16

    # ---- the dominant kernel inside the TIMED kind of step: GPU-side intervals of every kernel of a few more replays
    # of the same graph (CUPTI activity records), conv-busy time = union of the conv_igemm_kernel intervals -----------
    share = conv_busy = span = None
    by_name = {}
    try:
        calls = kernel_intervals(step_device, 5, flush)
        conv_busy = sum(busy_us([(a, b) for n, a, b in c if "conv_igemm_kernel" in n]) for c in calls) / len(calls)
        span = sum(max(b for _, _, b in c) - min(a for _, a, _ in c) for c in calls) / len(calls)
        share = conv_busy / span
        for c in calls:
            for n, a, b in c:
                key = n.split("(")[0].split("<")[0].replace("void ", "").strip()[-60:]
                by_name[key] = by_name.get(key, 0.0) + (b - a) / len(calls)
    except Exception as err:                        # no CUPTI: fall back to the whole step as the denominator
        by_name = {"error": f"{type(err).__name__}: {err}"}
    clocks = sampler.stop() if rank == 0 else {}
    conv_fl = sum(f[0] for f in flop_log)
    conv_fl_eff = sum(f[1] for f in flop_log)
    head_fl = sum(f[0] for f in flop_log if f[2] == "head")
    n_conv = len(flop_log)

    dev_ms, e2e_ms, skip_ms = dist_max([dev_ms, e2e_ms, skip_ms], dev, world)

    extra = {}
    del eng_skip, eng_skip_eager
    torch.cuda.empty_cache()
    for leg in legs:
        try:
            if leg == "cityscapes_sweep":
                extra[leg] = leg_cityscapes_sweep(dev, rank, world, max(4, steps // 2), warmup, barrier)
            elif leg in ("train_ce", "train_lovasz"):
                extra[leg] = leg_train(leg.split("_")[1], dev, rank, world, max(4, steps // 2), warmup, barrier)
        except Exception as err:
            import traceback
            extra[leg] = {"error": f"{type(err).__name__}: {err}", "trace": traceback.format_exc()[-600:]}

    if rank == 0:
        peaks = measured_peaks()
        imgs = PER_GPU_BATCH * world * steps
        fh, fw = ((d - 1) // 8 + 1 for d in img_hw())
        assert head_fl == head_flops(fh, fw, PER_GPU_BATCH, [1024, 2048, 2048]), "head FLOP model out of date"
        ms_step = dev_ms / steps
        conv_ms = ms_step * share if share is not None else ms_step
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12
        achieved_eff = conv_fl_eff / (conv_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        res = exit_res
        top = sorted(((v, k) for k, v in by_name.items() if isinstance(v, float)), reverse=True)[:8]
        line = {
            "metric": METRIC, "value": imgs / (dev_ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": imgs / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": (Xh16.numel() * 2 + yh8.numel()) if eng.use_graph else (Xh.numel() * 4 + yh.numel() * 8),
                    "d2h_bytes_per_step": PER_GPU_BATCH * 4 + 2 * PER_GPU_BATCH * 4, "ms_per_step": e2e_ms / steps,
                    "h2d_ms_alone": h2d_ms,
                    "inputs": "pinned host batches: bf16 images [N,3,H,W] + uint8 labels (bit-identical exits, maps and "
                              "confusion matrices to the fp32 / int64 upload: checked in this run)" if eng.use_graph else "fp32 / int64",
                    "fp32_int64_inputs": {"value": imgs / (max(e2e32_ms, 1e-9) * 1e-3) if e2e32_ms else None,
                                          "h2d_bytes_per_step": Xh.numel() * 4 + yh.numel() * 8, "h2d_ms_alone": h2d32_ms,
                                          "note": "rank 0 timing"},
                    "mode": "pipelined: upload of batch k+1 and read-back of batch k-1 overlap the graph replay of batch k"
                            if eng.use_graph else "sequential",
                    "l2": "every step's inputs arrive from pinned host memory (never cache-resident); no flush between "
                          "batches, so weights may stay in L2 as in a real stream — `value` is the flushed number"},
            "gpu_launches": int(launches) if args.no_graph else int(launches_per_step * steps),
            "launch_mode": "eager" if args.no_graph else "cuda_graph_replay (eeseg kernels captured in the graph)",
            "roofline": {"kernel": f"conv_igemm_kernel (tcgen05 implicit GEMM, {n_conv} launches/step: "
                                   "exit heads + ResNet stem / bottlenecks)",
                         "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "frac_nominal": achieved / peak, "frac_effective": achieved_eff / peak,
                         "achieved_effective": achieved_eff, "traffic": conv_traffic(),
                         "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                         "conv_ms_per_step": conv_ms, "conv_share_of_step": share,
                         "flops_per_step": conv_fl, "flops_per_step_effective": conv_fl_eff,
                         "flops_exit_heads": head_fl,
                         "timing": "conv_ms_per_step = ms_per_step (CUDA events around the graph replays of the timed region) x "
                                   "conv share; share = union of the conv_igemm_kernel intervals / span of the step, from the "
                                   "GPU-side kernel timestamps (CUPTI activity records via torch.profiler) of 5 further replays of "
                                   "the SAME graph with the same L2 flush; nominal FLOPs count padding taps the kernel skips, "
                                   "effective FLOPs do not",
                         "step_span_us_profiled": span, "conv_busy_us_profiled": conv_busy,
                         "kernel_us_per_step_top": [{"kernel": k, "us": round(v, 1)} for v, k in top]},
            "clocks": clocks,
            "wall_ms_timed_region": wall_ms,
            "exit_stats": {k: res[k] for k in ("b1_count", "b2_count", "count_out", "out_gl")},
            "early_exit_operating_point": {
                "tau": tau_mid, "value": imgs / (skip_ms * 1e-3), "unit": "images/s", "ms_per_step": skip_ms / steps,
                "mode": "skip_compute engine, one CUDA graph per (exit stage, active image count), one 4-byte D2H per gate "
                        "for the active count; rank 0 counters" if eng_skip_use_graph(args) else "skip_compute engine, eager launches",
                "e2e": {"value": imgs / (skip_e2e_ms * 1e-3), "unit": "images/s",
                        "how": "pinned host batches (bf16 images, uint8 labels) through evaluate_pipelined, rank 0 wall clock (x world size)"} if skip_e2e_ms else None,
                "eager_launch_value": imgs / (skip_eager_ms * 1e-3),
                "images_per_exit": skip_counts[:-1],
                "pct_images_exited_early": 100.0 * sum(skip_counts[:-2]) / max(1, skip_counts[-1]),
                "pct_pixels_below_tau_per_gate": [100.0 * px / max(1, PER_GPU_BATCH * steps * img_hw()[0] * img_hw()[1])
                                                  for px in skip_px[:-1]],
            },
            "extra_workloads": extra,
        }
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_reference_run(steps=2, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def eng_skip_use_graph(args):
    return not args.no_graph


if __name__ == "__main__":
    sys.exit(main())
