"""Per-stage device time of the graphed compute-skipping engine (engine.py: one CUDA graph per exit stage and
active-image count): replays every (stage, n) graph on its own with CUDA events. Usage: python tools/skip_stage_times.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ee_semantic_segmentation_b200.engine import EarlyExitEngine  # noqa: E402
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3  # noqa: E402


def main():
    bench.set_workload(sys.argv[1] if len(sys.argv) > 1 else "voc513")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, bench.img_hw()[0], sections=bench.SECTIONS, pretrained=False,
                        num_classes=bench.N_CLASSES).to(dev).eval()
    X, y = bench.synth_batch(0, bench.PER_GPU_BATCH)
    X, y = X.to(dev), y.to(dev)
    eng = EarlyExitEngine(net, bench.N_CLASSES, -1.0, skip_compute=True, use_graph=True)   # tau < 0: nobody leaves
    eng.evaluate(X, y)
    st = eng._skip_state(tuple(X.shape), True)
    rows = {}
    N = X.shape[0]
    for i in range(eng.E):
        for n in range(1, N + 1):
            if i == 0 and n != N:
                continue
            g = eng._skip_graph(st, (i, n), lambda: eng._skip_stage(st, i, n))
            for _ in range(3):
                g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            rows[f"stage{i}_n{n}_us"] = round(a.elapsed_time(b) / 20 * 1e3, 1)
    g = eng._skip_graph(st, 'final', lambda: eng._skip_final(st))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    rows["final_us"] = round(a.elapsed_time(b) / 20 * 1e3, 1)
    print(json.dumps(rows))
    if "--profile" in sys.argv:
        from torch.profiler import ProfilerActivity, profile
        g = st['stages'][(1, N)]
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_time_total > 0), key=lambda e: e.time_range.start)
        per = len(evs) // 5
        t0 = evs[-per].time_range.start
        for e in evs[-per:]:
            print(f"{e.time_range.start - t0:9.1f} {e.device_time_total:8.1f}  {e.name[:90]}")


if __name__ == "__main__":
    main()
