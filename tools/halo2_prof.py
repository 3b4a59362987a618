#!/usr/bin/env python
"""Role-cycle counters of the CTA-pair halo conv kernel (profile build) on the big 3x3 layers.
    DYK_B200_LIB=double-yolo-kaist_b200/libdyk_b200_prof.so python tools/halo2_prof.py"""
import ctypes as C, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops, _native as nat
from dyk.ops import View
dt = torch.float16
for (N, Cin, H, W, Cout, res) in [(16, 32, 256, 320, 64, True), (16, 32, 256, 320, 64, False), (16, 64, 128, 160, 128, True), (16, 128, 64, 80, 256, True), (16, 256, 32, 40, 512, True), (16, 512, 16, 20, 1024, True),
                                  (16, 512, 32, 40, 512, False), (16, 256, 32, 40, 512, False)]:
    x = View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin)
    y = View(torch.empty((N, H, W, Cout), device="cuda", dtype=dt), 0, Cout)
    r = View(torch.randn((N, H, W, Cout), device="cuda").to(dt), 0, Cout) if res else None
    w = (torch.randn((Cout, 3, 3, Cin), device="cuda") / (Cin * 9) ** 0.5).to(dt)
    sc, bi = torch.ones(2048, device="cuda"), torch.zeros(2048, device="cuda")
    for _ in range(2):
        ops.nhwc_conv(x, w, sc, bi, y, k=3, stride=1, pad=1, act="leaky", res=r)
    prof = torch.zeros(16, dtype=torch.int64, device="cuda"); prof[8] = 1 << 62; prof[12] = 1 << 62
    nat.call("dyk_conv_set_profile", C.c_void_p(prof.data_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.nhwc_conv(x, w, sc, bi, y, k=3, stride=1, pad=1, act="leaky", res=r); b.record()
    torch.cuda.synchronize()
    nat.call("dyk_conv_set_profile", None)
    p = prof.tolist(); n = max(p[7], 1)
    subs = ((W + 7) // 8) * ((H + 15) // 16) * N
    tiles = (subs + 1) // 2 * ((Cout + 255) // 256)
    n = min(n, 74) if n > 74 else n
    tpc = tiles / n
    print(f"{Cin}->{Cout} {H}x{W} res={int(res)}: {a.elapsed_time(b) * 1e3:7.1f} us  clusters {n} tiles/cluster {tpc:.2f} | MMA loop {p[2] / n:8.0f} cyc "
          f"(wait data {p[0] / max(p[2], 1):4.0%}, wait acc {p[1] / max(p[2], 1):4.0%}) | epilogue total {p[4] / n:8.0f} cyc: wait acc {p[3] / max(p[4], 1):4.0%}, "
          f"tmem ld {p[5] / max(p[4], 1):4.0%}, store-buffer wait {p[6] / max(p[4], 1):4.0%}, busy/tile {(p[4] - p[3]) / n / tpc:7.0f} cyc")
    t0 = p[8]
    print(f"      timeline (globaltimer, us after the first CTA's entry): last entry {(p[9] - t0) / 1e3:5.1f} | last prologue end "
          f"{(p[10] - t0) / 1e3:5.1f} | role loops end: first CTA {(p[12] - t0) / 1e3:5.1f}, last CTA {(p[11] - t0) / 1e3:5.1f} | last exit "
          f"{(p[13] - t0) / 1e3:5.1f}")
