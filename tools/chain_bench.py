#!/usr/bin/env python
"""What a convolution launch costs INSIDE a CUDA graph: a chain of n dependent launches of one layer shape (ping-pong
buffers where the shapes allow, else the same input) captured in a graph with programmatic dependent launch, replayed and
timed; beside it the eager event bracket of one launch and, with the profile build, the in-kernel globaltimer timeline.
    [DYK_B200_LIB=.../libdyk_b200_prof.so] python tools/chain_bench.py"""
import ctypes as C, os, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops, _native as nat
from dyk.ops import View
PROFILE = os.environ.get("DYK_B200_LIB", "").endswith("_prof.so")
dt = torch.float16
SHAPES = [  # (N, Cin, H, W, Cout, k, res)
    (16, 256, 64, 80, 128, 1, False), (16, 512, 32, 40, 256, 1, False), (16, 1024, 16, 20, 512, 1, False),
    (16, 128, 64, 80, 256, 3, True), (16, 256, 32, 40, 512, 3, True), (16, 512, 16, 20, 1024, 3, True),
    (16, 64, 256, 320, 32, 1, False), (16, 32, 256, 320, 64, 3, True), (16, 64, 128, 160, 128, 3, True),
    # CSPDarknet53 (dyolov4_fshare) residual stages
    (16, 128, 64, 80, 128, 3, True), (16, 256, 32, 40, 256, 3, True), (16, 512, 16, 20, 512, 3, True),
    (16, 64, 256, 320, 64, 1, False), (16, 128, 128, 160, 128, 1, False), (16, 256, 64, 80, 256, 1, False),
]
SHAPES += [(16, 256, 64, 80, 256, 3, False), (16, 512, 32, 40, 512, 3, False), (16, 1024, 16, 20, 1024, 3, False)]   # 15-17: the fusion convs
DUAL = os.environ.get("DYK_CHAIN_DUAL") == "1"      # run them with the dual-source operand (fused modality fusion)
if os.environ.get("DYK_CHAIN_ONLY"):
    SHAPES = [SHAPES[int(i)] for i in os.environ["DYK_CHAIN_ONLY"].split(",")]
n = 40
for (N, Cin, H, W, Cout, k, res) in SHAPES:
    x = View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin)
    ys = [View(torch.empty((N, H, W, Cout), device="cuda", dtype=dt), 0, Cout) for _ in range(2)]
    r = View(torch.randn((N, H, W, Cout), device="cuda").to(dt), 0, Cout) if res else None
    w = (torch.randn((Cout, k, k, Cin), device="cuda") / (Cin * k * k) ** 0.5).to(dt)
    sc, bi = torch.ones(2048, device="cuda"), torch.zeros(2048, device="cuda")
    kw = {}
    if DUAL and k == 3 and not res and Cout >= 256:
        kw = dict(x2=View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin), x_wts_raw=torch.tensor([0.3, -0.2], device="cuda"))
    run = lambda i: ops.nhwc_conv(x, w, sc, bi, ys[i & 1], k=k, stride=1, pad=k // 2, act="leaky", res=r, **kw)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    ev = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(i); b.record(); torch.cuda.synchronize()
        ev.append(a.elapsed_time(b) * 1e3)
    eager = sorted(ev)[len(ev) // 2]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            run(i)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / n)
    chain = sorted(ts)[len(ts) // 2]
    fl = 2.0 * N * H * W * Cout * Cin * k * k
    by = 2.0 * (N * H * W * Cin + N * H * W * Cout * (2 if res else 1) + Cout * Cin * k * k)
    roof = max(fl / 1.59e15, by / 6.4e12) * 1e6
    line = f"{Cin:5d}->{Cout:5d} k{k} {H}x{W} res={int(res)}: eager bracket {eager:6.1f} us | in-graph chain {chain:6.1f} us/launch | roofline {roof:5.1f} us"
    if PROFILE:
        prof = torch.zeros(16, dtype=torch.int64, device="cuda"); prof[8] = 1 << 62; prof[12] = 1 << 62
        nat.call("dyk_conv_set_profile", C.c_void_p(prof.data_ptr()))
        run(0); torch.cuda.synchronize()
        nat.call("dyk_conv_set_profile", None)
        p = prof.tolist(); t0 = p[8]
        line += (f" | timeline us: last entry {(p[9] - t0) / 1e3:4.1f}, last prologue end {(p[10] - t0) / 1e3:4.1f}, loops end first "
                 f"{(p[12] - t0) / 1e3:5.1f} last {(p[11] - t0) / 1e3:5.1f}, last exit {(p[13] - t0) / 1e3:5.1f}")
    print(line, flush=True)
