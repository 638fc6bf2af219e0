#!/usr/bin/env python
"""Per-role wait-time breakdown of the persistent conv kernel for a few layer shapes (tuning aid)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ee_semantic_segmentation_b200 import _lib
from ee_semantic_segmentation_b200.head_plan import conv_igemm

if not hasattr(_lib.lib(), "eeseg_conv_debug_stats"):
    raise SystemExit("this tool needs the tuning build: python -m ee_semantic_segmentation_b200.build --tuning, then run with "
                     "EESEG_LIB=ee_semantic_segmentation_b200/libeeseg_b200_tuning.so (the product library has no EESEG_TUNING hooks)")
dev = torch.device("cuda:0")
if os.environ.get("EESEG_PROBE"):   # operand-stream probe (see tools/conv_stream_probe.py): e.g. 15 = no loads, no stores
    _lib.lib().eeseg_conv_probe(int(os.environ["EESEG_PROBE"]), 0)
shapes = [  # name, N, h, w, Cin, Cout, R, dil, stride, residual
    ("l1.c3 64>256+res", 4, 129, 129, 64, 256, 1, 1, 1, True),
    ("l1.c1 256>64", 4, 129, 129, 256, 64, 1, 1, 1, False),
    ("l1.c2 3x3 64", 4, 129, 129, 64, 64, 3, 1, 1, False),
    ("l3.c3 256>1024+res", 4, 65, 65, 256, 1024, 1, 1, 1, True),
    ("l3.c1 1024>256", 4, 65, 65, 1024, 256, 1, 1, 1, False),
    ("l3.c2 3x3 256 d2", 4, 65, 65, 256, 256, 3, 2, 1, False),
    ("l4.c3 512>2048+res", 4, 65, 65, 512, 2048, 1, 1, 1, True),
    ("l4.c2 3x3 512 d4", 4, 65, 65, 512, 512, 3, 4, 1, False),
    ("aspp 3x3 d12 2048", 4, 65, 65, 2048, 256, 3, 12, 1, False),
]
buf = torch.zeros(148, 32, dtype=torch.int64, device=dev)
for name, N, h, w, cin, cout, R, dil, stride, res in shapes:
    x = torch.randn(N, h, w, cin, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, R, R, cin, device=dev) * 0.02).to(torch.bfloat16)
    sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    out = torch.empty(N, h, w, cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(N, h, w, cout, device=dev).to(torch.bfloat16) if res else None
    run = lambda: conv_igemm(x, wt, None, sh, dil, True, out, _lib.BF16, cout, stride=stride, residual=r)   # scale folded (NULL), as the inference plans launch it
    for _ in range(3):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10):
        run()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    buf.zero_()
    _lib.lib().eeseg_conv_debug_stats(buf.data_ptr())
    run(); torch.cuda.synchronize()
    _lib.lib().eeseg_conv_debug_stats(None)
    d = buf.cpu().double()
    act = d[:, 3] > 0
    m = d[act].mean(0)
    tiles = m[3]
    di = buf.cpu()
    kern_us = (di[act][:, 17].max() - di[act][:, 16].min()).item() / 1e3
    pro_us = ((di[act][:, 17] - di[act][:, 16]).double().mean().item() - m[15].item()) / 1e3
    print(f"{name:22s} event {us:6.1f}us KERNEL {kern_us:6.1f}us (outside epilogue role {pro_us:4.1f}us) tiles/CTA {tiles:4.1f} | PROD total {m[2]:8.0f} wait_res_empty {m[0]:7.0f} wait_empty {m[1]:7.0f} | "
          f"MMA total {m[6]:8.0f} wait_tmem_empty {m[4]:7.0f} wait_full {m[5]:7.0f} | "
          f"EPI total {m[12]:8.0f} wait_tmem_full {m[8]:7.0f} wait_res {m[9]:7.0f} wait_store_read {m[10]:7.0f} bar1 {m[11]:6.0f} ss_load {m[13]:6.0f} colloop {m[14]:7.0f} fence_bar2 {m[18]:6.0f} store_issue {m[19]:6.0f} epi_ns {m[15]:7.0f} => {m[12]/max(m[15],1):.2f} GHz", flush=True)

# ---- grouped ASPP launch vs the four separate launches (kernel wall time from the dbg stamps) ----
from ee_semantic_segmentation_b200.head_plan import conv_igemm_grouped, group_schedule
for cin in (2048, 1024):
    N, h, w, mid = 4, 65, 65, 256
    x = torch.randn(N, h, w, cin, device=dev).to(torch.bfloat16)
    ks, dl = [1, 3, 3, 3], [1, 12, 24, 36]
    wts = [(torch.randn(mid, k, k, cin, device=dev) * 0.02).to(torch.bfloat16) for k in ks]
    scs = [torch.ones(mid, device=dev) for _ in ks]
    shs = [torch.zeros(mid, device=dev) for _ in ks]
    cat = torch.empty(N, h, w, 4 * mid, dtype=torch.bfloat16, device=dev)
    sched = group_schedule(N, h, w, cin, mid, ks, dl).to(dev)

    def stamp(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        buf.zero_()
        _lib.lib().eeseg_conv_debug_stats(buf.data_ptr())
        fn(); torch.cuda.synchronize()
        _lib.lib().eeseg_conv_debug_stats(None)
        di = buf.cpu()
        act = di[:, 17] > 0
        return (di[act][:, 17].max() - di[act][:, 16].min()).item() / 1e3, di

    tot = 0.0
    for k in range(4):
        us, _ = stamp(lambda: conv_igemm(x, wts[k], scs[k], shs[k], dl[k], True, cat[..., k * mid:], _lib.BF16, 4 * mid))
        print(f"  Cin={cin} separate k={ks[k]} d={dl[k]}: {us:.1f} us")
        tot += us
    n_cl = _lib.lib().eeseg_conv_pair_clusters()
    psched = group_schedule(N, h, w, cin, mid, ks, dl, pairs=True, n_clusters=n_cl).to(dev)
    for label, sc_, pr_ in (("GROUPED", sched, False), (f"GROUPED as CTA pairs ({n_cl} clusters)", psched, True)):
        us, di = stamp(lambda: conv_igemm_grouped(x, wts, scs, shs, ks, dl, [k * mid for k in range(4)], True, cat, 4 * mid,
                                                  4 * mid, sc_, cta_pairs=pr_))
        d = di.double()
        act = d[:, 6] > 0          # CTAs whose MMA thread ran (pair launches: the leaders)
        m = d[act].mean(0)
        ep = d[d[:, 12] > 0]
        print(f"  Cin={cin} separate total {tot:.1f} us | {label} {us:.1f} us  (MMA total {m[6]:.0f} cyc, wait_full {m[5]:.0f}, "
              f"wait_tmem_empty {m[4]:.0f}; per-CTA MMA total min/max {d[act][:,6].min():.0f}/{d[act][:,6].max():.0f}; "
              f"EPI total mean {ep[:,12].mean():.0f} wait_tmem_full {ep[:,8].mean():.0f} colloop {ep[:,14].mean():.0f}; "
              f"PROD total {d[d[:,2]>0][:,2].mean():.0f} wait_empty {d[d[:,2]>0][:,1].mean():.0f})")
        run = di[di[:, 17] > 0]
        st, en = run[:, 16], run[:, 17]
        print(f"      CTA start skew {(st.max() - st.min()).item() / 1e3:.1f} us, end skew {(en.max() - en.min()).item() / 1e3:.1f} us, "
              f"CTA lifetime min/mean/max {(en - st).min().item() / 1e3:.1f}/{(en - st).double().mean().item() / 1e3:.1f}/{(en - st).max().item() / 1e3:.1f} us, "
              f"MMA role mean {m[6] / 1e3:.0f} k cycles, epilogue role wall {ep[:, 15].mean() / 1e3:.1f} us (=> {ep[:, 12].mean() / ep[:, 15].mean():.2f} GHz)")
