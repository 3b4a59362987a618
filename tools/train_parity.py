#!/usr/bin/env python
"""Gradient parity of the native training path against the oracle's fp32 autograd on one seeded step.
    python tools/train_parity.py [cfg name | path.cfg] [H] [W] [B] [fp16|bf16]
Loss = sum_i <p_i, R_i> with fixed random R_i (a linear probe of every head output), so d loss / d p is known and
identical on both sides.  Prints the relative L2 error of every parameter gradient and of the BN running stats."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import models
from dyk import cfg_zoo
from oracle import darknet_ref as dr


from oracle.train_check import run, LAST


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else str(REPO / "tests/data/tiny_yolov3_train.cfg")
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    dt = torch.bfloat16 if (len(sys.argv) > 5 and sys.argv[5] == "bf16") else torch.float16
    rows, stats, fwd, loss, loss_o = run(cfg, H, W, B, dt)
    print(f"{cfg} {H}x{W} B={B} {dt}: loss native {loss:.6g} oracle {loss_o:.6g}; head rel err {['%.2e' % e for e in fwd]}")
    for name, rel, den in rows:
        print(f"{name:60s} rel_l2 = {'missing' if rel is None else '%.3e' % rel}   |grad| = {'-' if den is None else '%.3e' % den}")
    rels = sorted(r for _, r, _ in rows if r is not None)
    print(f"params {len(rows)}: median rel err {rels[len(rels) // 2]:.3e}, max {rels[-1]:.3e}, missing {sum(r is None for _, r, _ in rows)}")
    print("running stats: max rel err", max(e for _, e in stats))

    # teacher-forced, block-by-block backward check (tight): see oracle/layerwise.compare_backward
    from oracle import layerwise
    import models as _m
    # (re-run one step so the plan buffers hold exactly this step's tensors)
    print("--- teacher-forced backward, per convolution block (relative L2 error)")
    worst = {}
    for r in layerwise.compare_backward(LAST["plan"], LAST["frames"]):
        print("L%-4d " % r["layer"] + "  ".join(f"{k}={v:.2e}" for k, v in r.items() if k != "layer"))
        for k, v in r.items():
            if k != "layer":
                worst[k] = max(worst.get(k, 0.0), v)
    print("worst:", {k: "%.2e" % v for k, v in worst.items()})
