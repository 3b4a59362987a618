"""Host-side logic of the plan compilers (no GPU): SPP pool cascade, and the schedule that tells the overlapped gradient
all-reduce which ranges of the flat gradient buffer are final after each backward step."""
import pytest
import torch

from dyk import cfg_zoo


def _model(name, size=(64, 96)):
    import models
    return models.YOLO(cfg_zoo.materialize(name), size)


def test_spp_pools_become_a_cascade_of_5x5():
    from dyk import plan as P
    m = _model("kaist_dyolov4_fshare_global_concat_se3.cfg").eval()
    ops_, _, _, _ = P.build_ops(m, 64, 96, True)
    ops_ = P.fuse(ops_)
    pools = [o for o in ops_ if isinstance(o, P.PoolOp)]
    assert sorted(o.k for o in pools) == [5, 9, 13] and len({id(o.src) for o in pools}) == 1
    x = pools[0].src
    uses_before = x.uses
    P.cascade_pools(ops_)
    assert [o.k for o in pools] == [5, 5, 5]
    assert pools[0].src is x and pools[1].src is pools[0].out and pools[2].src is pools[1].out
    assert x.uses == uses_before - 2 and pools[0].out.uses >= 2      # the concat and the next pool


@pytest.mark.parametrize("name,dual", [("kaist_dyolov4_fshare_global_concat_se3.cfg", True),
                                       ("kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg", True),
                                       ("kaist_yolov3.cfg", False)])
def test_gradient_ready_schedule_covers_the_flat_buffer_once(name, dual):
    from dyk import dist_utils as du
    from dyk import train_plan as TP
    m = _model(name).train()
    plan = TP.TrainPlan(m, 2, 64, 96, torch.bfloat16, dual, torch.device("cpu"))
    ranges = sorted(r for v in plan.ready_after.values() for r in v)
    assert ranges[0][0] == 0 and ranges[-1][1] == plan.grad_numel
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])), "gaps or overlaps in the flat gradient buffer"
    assert max(plan.ready_after) < len(plan.bwd)

    class Recorder(du.OverlappedAllReduce):
        def _active(self):
            return True

        def _native_avg(self):
            return False

        def _send(self, lo, hi):
            self.sent.append((lo, hi))

    red = Recorder(bucket_bytes=8 << 20)
    red.sent = []
    red.begin(None, plan.grad_numel)
    red.feed(plan.ready_after.get(-1, ()))
    early = 0
    for i in range(len(plan.bwd)):
        red.feed(plan.ready_after.get(i, ()))
        if i == len(plan.bwd) // 2:
            early = len(red.sent)
    red._flush(final=True)
    assert not red.ready and red.sent_lo == 0
    assert all(a[0] == b[1] for a, b in zip(red.sent, red.sent[1:])), "buckets must be contiguous, from the end downwards"
    assert sum(hi - lo for lo, hi in red.sent) == plan.grad_numel
    if plan.grad_numel * 4 > 4 * (8 << 20):
        assert early >= 1, "no bucket was ready half-way through the backward pass"


def _eval_ops(name, H=512, W=640, B=16, dtype=torch.float16):
    from dyk import plan as P
    m = _model(name, (H, W)).eval()
    dual = "second_index" in m.net_info
    raw, vals, _, _ = P.build_ops(m, H, W, dual)
    ops_ = P.fuse_se_gates(P.fuse_weighted_adds(P.fuse(raw), B, dtype), B, dtype)
    return P, ops_, vals


def test_modality_fusion_moves_into_the_consuming_conv(native_lib):
    """dyolov3_add_sl (the bench model): the three weighted [shortcut]s (visible + LWIR at strides 8 / 16 / 32,
    layers.py:63-85) have a 3x3 stride-1 consumer each and disappear into it (dual-source operand); dyolov4_fshare's four
    feed stride-2 convolutions and stay stand-alone."""
    P, ops_, vals = _eval_ops("kaist_dyolov3_add_sl.cfg")
    assert not [o for o in ops_ if isinstance(o, P.AddOp) and o.module.weight]
    duals = [o for o in ops_ if isinstance(o, P.ConvOp) and o.src2 is not None]
    assert [(o.src.C, o.src.H) for o in duals] == [(256, 64), (512, 32), (1024, 16)]
    assert all(o.fusion.weight and o.conv.kernel_size == (3, 3) for o in duals)
    assert sum(1 for v in vals if v is not None and v.virtual) == 3          # the three fused sums own no buffer
    P, ops_, _ = _eval_ops("kaist_dyolov4_fshare_global_concat_se3.cfg")
    assert len([o for o in ops_ if isinstance(o, P.AddOp) and o.module.weight]) == 4
    # small maps (test sizes) fall below the sub-tile efficiency threshold: nothing is fused, nothing breaks
    P, ops_, _ = _eval_ops("kaist_dyolov3_add_sl.cfg", 64, 96, 2)
    assert len([o for o in ops_ if isinstance(o, P.AddOp) and o.module.weight]) + \
        len([o for o in ops_ if isinstance(o, P.ConvOp) and o.src2 is not None]) == 3


@pytest.mark.parametrize("name,blocks", [("kaist_dyolov4_fshare_global_concat_se3.cfg", 3),
                                         ("kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg", 19)])
def test_se_blocks_hand_their_gate_to_the_consumers(native_lib, name, blocks):
    """Every [se] block of the two BASELINE models that have them computes only its gate (layers.py:184-189); the
    `scale * x` (:190) is applied by the consumers, which read the block's input."""
    P, ops_, _ = _eval_ops(name)
    ses = [o for o in ops_ if isinstance(o, P.SEOp)]
    assert len(ses) == blocks and all(o.gate_only and o.out.virtual for o in ses)
    for se in ses:
        cons = [o for o in ops_ if getattr(o, "gate_of", None) is se]
        assert cons, "a gate-only block without a consumer"
        for c in cons:
            if isinstance(c, P.ConvOp):
                assert c.src is se.src and c.conv.kernel_size == (1, 1)
            else:
                assert isinstance(c, P.AddOp) and c.x is se.src
        assert not [o for o in ops_ if se.out in o.inputs() and getattr(o, "gate_of", None) is not se]
