"""Drop-in proof: the reference's own callers of the hot path, restated, driving the native module API.

`train_one_epoch` (reference train_utils/kaist_train_eval_utils.py:12-118) and `evaluate` (:121-190) are the two functions
that sit directly on top of models.YOLO / compute_loss / non_max_suppression / scale_coords.  /root/reference does not
travel to the GPU box, and its train_utils package does not import on this image anyway (torch._six, pycocotools —
SURVEY.md App. F), so the two loop bodies are restated here statement by statement (same call sequence, same arguments:
uint8 -> float / 255, multi-scale F.interpolate, amp.autocast + GradScaler, loss accumulation, scaler.step / update /
zero_grad every `accumulate` batches; eval: model(v, l)[0] -> NMS(0.01, 0.6, multi_label=False) -> scale_coords().round())
with the logging / COCO bookkeeping removed, and run against `models`, `build_utils.utils` of THIS repo — the modules a
maintainer drops over the reference's.  Both optimizers of train.py:85-91 are exercised, stock torch.optim
and the fused dyk.optim ones."""
import math
import random

import pytest
import torch
import torch.nn.functional as F
from torch.cuda import amp

from dyk import cfg_zoo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
H, W = 128, 160


class SyntheticPairs:
    """What the reference's DataLoader yields per batch (kaist_dataset.py collate): uint8 visible / LWIR frames,
    targets [img, cls, x, y, w, h] (normalised), paths, shapes ((h0, w0), ((h/h0, w/w0), pad)), image indices."""

    def __init__(self, batches, bs, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for b in range(batches):
            v = torch.randint(0, 256, (bs, 3, H, W), dtype=torch.uint8, generator=g)
            l = torch.randint(0, 256, (bs, 3, H, W), dtype=torch.uint8, generator=g)
            t = torch.zeros((2 * bs, 6))
            t[:, 0] = torch.arange(2 * bs) // 2
            t[:, 2:4] = torch.rand((2 * bs, 2), generator=g) * 0.6 + 0.2
            t[:, 4:6] = torch.rand((2 * bs, 2), generator=g) * 0.2 + 0.1
            shapes = [((2 * H, 2 * W), ((0.5, 0.5), (0.0, 0.0))) for _ in range(bs)]     # originals were 2x larger
            self.items.append((v, l, t, [f"img{b}_{i}" for i in range(bs)], shapes, list(range(b * bs, (b + 1) * bs))))

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


def train_one_epoch(model, optimizer, dataloader, device, epoch, accumulate, img_size, grid_min, grid_max, gs, multi_scale,
                    compute_loss, fused_input=False):
    """kaist_train_eval_utils.py:12-118 without MetricLogger / reduce_dict / warm-up scheduler.
    fused_input=True is the integration INTEGRATION.md recommends: the uint8 frames go to the model as they are (`/ 255` in
    the stem) and the multi-scale resize becomes `model(v, l, input_size=ns)` instead of two F.interpolate calls."""
    model.train()
    enable_amp = "cuda" in device.type                                        # :41
    scaler = amp.GradScaler(enabled=enable_amp)                               # :42
    mloss = torch.zeros(4).to(device)
    nb = len(dataloader)
    history = []
    for i, (v_imgs, l_imgs, targets, paths, _, _) in enumerate(dataloader):
        ni = i + nb * epoch
        if fused_input:
            v_imgs, l_imgs = v_imgs.to(device), l_imgs.to(device)
        else:
            v_imgs = v_imgs.to(device).float() / 255.0                        # :54-55
            l_imgs = l_imgs.to(device).float() / 255.0
        targets = targets.to(device)
        ns = None
        if multi_scale:                                                       # :59-71
            if ni % accumulate == 0:
                img_size = random.randrange(grid_min, grid_max + 1) * gs
            sf = img_size / max(v_imgs.shape[2:])
            if sf != 1:
                ns = [math.ceil(x * sf / gs) * gs for x in v_imgs.shape[2:]]
                if not fused_input:
                    v_imgs = F.interpolate(v_imgs, size=ns, mode='bilinear', align_corners=False)
                    l_imgs = F.interpolate(l_imgs, size=ns, mode='bilinear', align_corners=False)
        with amp.autocast(enabled=enable_amp):                                # :74
            pred = model(v_imgs, l_imgs, input_size=ns) if fused_input else model(v_imgs, l_imgs)
            loss_dict = compute_loss(pred, targets, model)
            losses = sum(loss for loss in loss_dict.values())
            loss_items = torch.cat((loss_dict["box_loss"], loss_dict["obj_loss"], loss_dict["class_loss"], losses)).detach()
            mloss = (mloss * i + loss_items) / (i + 1)
            assert torch.isfinite(losses)                                     # :95-98 (the reference exits)
            losses *= 1. / accumulate                                         # :100
        scaler.scale(losses).backward()                                       # :103
        if ni % accumulate == 0:                                              # :105-108
            scaler.step(optimizer)
            scaler.update()
            optimizer.zero_grad()
        history.append(float(loss_items[3]))
    train_one_epoch.last_scale = scaler.get_scale()
    return mloss, history


@torch.no_grad()
def evaluate(model, dataloader, device, non_max_suppression, scale_coords):
    """kaist_train_eval_utils.py:121-190 up to the per-image result dicts handed to the COCO evaluator."""
    cpu_device = torch.device("cpu")
    model.eval()
    res = {}
    for v_imgs, l_imgs, targets, paths, shapes, img_index in dataloader:
        v_imgs = v_imgs.to(device).float() / 255.0                            # :138-139
        l_imgs = l_imgs.to(device).float() / 255.0
        torch.cuda.synchronize(device)
        pred = model(v_imgs, l_imgs)[0]                                       # :148
        pred = non_max_suppression(pred, conf_thres=0.01, iou_thres=0.6, multi_label=False)   # :149
        outputs = []
        for index, p in enumerate(pred):
            if p is None:
                p = torch.empty((0, 6), device=cpu_device)
                boxes = torch.empty((0, 4), device=cpu_device)
            else:
                boxes = p[:, :4]
                boxes = scale_coords(v_imgs[index].shape[1:], boxes, shapes[index][0]).round()   # :163
            outputs.append({"boxes": boxes.to(cpu_device), "labels": p[:, 5].to(device=cpu_device, dtype=torch.int64),
                            "scores": p[:, 4].to(cpu_device)})
        res.update({img_id: output for img_id, output in zip(img_index, outputs)})
    return res


def _model(name, seed=0):
    import models
    torch.manual_seed(seed)
    m = models.YOLO(cfg_zoo.materialize(name), (H, W)).to(DEV)
    m.nc, m.gr = 1, 1.0
    m.hyp = {"box": 3.54, "cls": 37.4, "obj": 64.3, "cls_pw": 1.0, "obj_pw": 1.0, "iou_t": 0.20, "fl_gamma": 0.0}
    if "yolov4" in m.cfg:
        m.hyp["ciou"] = 1.0
    return m


@pytest.mark.parametrize("name", ["kaist_dyolov4_fshare_global_concat_se3.cfg", "kaist_dyolov3_add_sl.cfg"])
@pytest.mark.parametrize("opt_kind", ["torch_sgd", "fused_sgd", "fused_adam"])
def test_reference_training_loop_over_native_modules(native_lib, name, opt_kind):
    from build_utils.utils import compute_loss
    from dyk import optim
    random.seed(0)
    model = _model(name)
    pg = [p for p in model.parameters() if p.requires_grad]                   # train.py:85-91
    if opt_kind == "torch_sgd":
        optimizer = torch.optim.SGD(pg, lr=1e-3, momentum=0.937, weight_decay=5e-4, nesterov=True)
    elif opt_kind == "fused_sgd":
        optimizer = optim.FusedSGD(pg, lr=1e-3, momentum=0.937, weight_decay=5e-4, nesterov=True)
    else:
        optimizer = optim.FusedAdam(pg, lr=1e-3, betas=(0.937, 0.999), weight_decay=5e-4)
    before = [p.detach().clone() for p in pg[:8]]
    data = SyntheticPairs(batches=16, bs=2)      # 8 optimizer steps: room for GradScaler to back off from 65536 under fp16
    # multi-scale on: sizes 96 / 128 / 160 (train.py:143-151 picks grid_min / grid_max around img_size; gs = 32)
    mloss, hist = train_one_epoch(model, optimizer, data, DEV, epoch=0, accumulate=2, img_size=W, grid_min=3, grid_max=5,
                                  gs=32, multi_scale=True, compute_loss=compute_loss)
    torch.cuda.synchronize()
    if opt_kind == "fused_sgd":
        # the same epoch with the fused input path (uint8 frames, resize inside the stem) follows the same loss trajectory
        random.seed(0)
        model2 = _model(name)
        opt2 = optim.FusedSGD([p for p in model2.parameters() if p.requires_grad], lr=1e-3, momentum=0.937, weight_decay=5e-4,
                              nesterov=True)
        _, hist2 = train_one_epoch(model2, opt2, SyntheticPairs(batches=16, bs=2), DEV, epoch=0, accumulate=2, img_size=W,
                                   grid_min=3, grid_max=5, gs=32, multi_scale=True, compute_loss=compute_loss, fused_input=True)
        assert all(math.isfinite(h) for h in hist2)
        assert abs(hist2[0] - hist[0]) < 0.02 * abs(hist[0]), (hist[0], hist2[0])
    assert all(math.isfinite(h) for h in hist) and bool(torch.isfinite(mloss).all())
    assert train_one_epoch.last_scale >= 1.0, "GradScaler collapsed: the backward pass produces non-finite gradients at any scale"
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, pg[:8])), \
        f"the optimizer never changed the parameters (final loss scale {train_one_epoch.last_scale})"
    assert all(p.grad is None or bool(torch.isfinite(p.grad).all()) for p in pg)
    # same frames, fixed scale: a second short epoch keeps the loss finite and the BatchNorm statistics have moved
    bn = next(m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d))
    assert int(bn.num_batches_tracked) == len(data)


def test_reference_evaluate_loop_over_native_modules(native_lib):
    from build_utils.utils import non_max_suppression, scale_coords
    from oracle import darknet_ref as dr
    from oracle import weights as ow
    name = "kaist_dyolov3_add_sl.cfg"
    model = _model(name)
    ref = dr.DarknetRef(cfg_zoo.materialize(name))
    st = ow.make_calibrated_state(ref, seed=0)
    model.load_state_dict(st, strict=True)
    data = SyntheticPairs(batches=2, bs=3, seed=5)
    res = evaluate(model, data, DEV, non_max_suppression, scale_coords)
    assert sorted(res) == list(range(6))
    n_det = 0
    for out in res.values():
        b, s, lab = out["boxes"], out["scores"], out["labels"]
        assert b.shape[1:] == (4,) and b.shape[0] == s.shape[0] == lab.shape[0] <= 100
        if b.shape[0]:
            n_det += b.shape[0]
            assert bool((b == b.round()).all())
            assert float(b[:, [0, 2]].min()) >= 0 and float(b[:, [0, 2]].max()) <= 2 * W       # clipped to the ORIGINAL image
            assert float(b[:, [1, 3]].min()) >= 0 and float(b[:, [1, 3]].max()) <= 2 * H
            assert bool((s[:-1] >= s[1:]).all()) and float(s.min()) > 0.01 and bool((lab == 0).all())
    assert n_det > 0
