"""``.cfg``-driven Darknet model API — drop-in for the reference's models.py.

Same public names and semantics (``create_modules``, ``YOLOLayer``, ``YOLO``, ``load_darknet_weights``,
plus the star-import re-exports train.py relies on), same module tree and ``state_dict`` keys
(``module_list.{i}.Conv2d.weight`` ...).  What differs is execution: ``YOLO.forward`` does not walk the
module list op by op; it compiles the block list into an execution plan of fused sm_100a kernels
(dyk/plan.py) the first time a (mode, batch, H, W, dtype) combination is seen and replays it, by default
as a CUDA graph.  The ``nn`` modules only own the parameters.

Reference citations: models.py:7-155 (create_modules), :158-258 (YOLOLayer), :261-315 (YOLO),
:318-364 (load_darknet_weights).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from build_utils import torch_utils
from build_utils.layers import *  # noqa: F401,F403  (re-exported like the reference does)
from build_utils.layers import (DepthwiseSeparableConv2d, FeatureConcat, Inception, SqueezeExcitation,
                                WeightedFeatureFusion, _activation_module)
from build_utils.parse_config import *  # noqa: F401,F403
from build_utils.parse_config import parse_model_cfg
from build_utils.utils import get_yolo_layers
from dyk import ops as _ops
from dyk import plan as _plan
from dyk import train_plan as _train_plan


def create_modules(modules_defs: list, img_size, cfg):
    """Block dicts -> (nn.ModuleList, routs_binary, net_info).  Pops the [net] block from the list it is
    given, exactly like reference models.py:17-18."""
    img_size = [img_size] * 2 if isinstance(img_size, int) else img_size
    net_infos = modules_defs.pop(0)
    out_filters = [3]
    module_list = nn.ModuleList()
    routs = []
    yolo_index = -1
    second = net_infos.get("second_index", None)

    for i, mdef in enumerate(modules_defs):
        kind = mdef['type']
        modules = nn.Sequential()
        filters = out_filters[-1] if kind not in ('convolutional', 'depthwiseconvolutional', 'route') else None

        if kind == 'convolutional':
            bn = mdef['batch_normalize']
            filters = mdef['filters']
            k = mdef['size']
            stride = mdef['stride'] if 'stride' in mdef else (mdef['stride_y'], mdef['stride_x'])
            if not isinstance(k, int):
                raise TypeError("conv2d filter size must be int type")
            cin = 3 if (second is not None and i == second) else out_filters[-1]
            modules.add_module("Conv2d", nn.Conv2d(in_channels=cin, out_channels=filters, kernel_size=k,
                                                   stride=stride, padding=k // 2 if mdef['pad'] else 0,
                                                   groups=mdef['groups'] if 'groups' in mdef else 1,
                                                   bias=not bn))
            if bn:
                modules.add_module("BatchNorm2d", nn.BatchNorm2d(filters))
            else:
                routs.append(i)  # detection-head convs are always kept (models.py:49)
            act = _activation_module(mdef['activation'])
            if act is not None:
                modules.add_module("activation", act)

        elif kind == 'depthwiseconvolutional':
            filters = mdef['filters']
            stride = mdef['stride'] if 'stride' in mdef else (mdef['stride_y'], mdef['stride_x'])
            modules = DepthwiseSeparableConv2d(in_channels=out_filters[-1], out_channels=filters,
                                               kernel_size=mdef['size'] if 'size' in mdef else 3, stride=stride)

        elif kind == 'dropout':
            modules = nn.Dropout(mdef['probability'])

        elif kind == 'inception':
            modules = Inception(in_channels=out_filters[-1], n1x1=mdef['n1x1'], n3x3_reduce=mdef['n3x3_reduce'],
                                n3x3=mdef['n3x3'], n5x5_reduce=mdef['n5x5_reduce'], n5x5=mdef['n5x5'],
                                pool_proj=mdef['pool_proj'])
            # NB: like the reference (models.py:81-85) `filters` is left at the previous value here.

        elif kind == 'se':
            modules = SqueezeExcitation(in_channels=out_filters[-1], squeeze_factor=mdef['squeeze_factor'])

        elif kind == 'maxpool':
            k = mdef['size']
            modules = nn.MaxPool2d(kernel_size=k, stride=mdef['stride'], padding=(k - 1) // 2)

        elif kind == 'avgpool':
            modules = nn.AdaptiveAvgPool2d(output_size=mdef['size'])

        elif kind == 'upsample':
            modules = nn.Upsample(scale_factor=mdef['stride'])

        elif kind == 'route':
            layers = mdef['layers']
            filters = sum(out_filters[l + 1 if l > 0 else l] for l in layers)
            layers = [i + l if l < 0 else l for l in layers]
            routs.extend(layers)
            modules = FeatureConcat(layers=layers)

        elif kind == 'shortcut':
            layers = [i + l if l < 0 else l for l in mdef['from']]
            routs.extend(layers)
            modules = WeightedFeatureFusion(layers=layers, weight='weights_type' in mdef)

        elif kind == 'yolo':
            yolo_index += 1
            stride = [8, 16, 32, 64, 128]
            if any(x in cfg for x in ['yolov-tiny', 'fpn', 'yolov3']):
                stride = [32, 16, 8]
            modules = YOLOLayer(anchors=mdef['anchors'][mdef['mask']], nc=mdef['classes'], img_size=img_size,
                                stride=stride[yolo_index], bf_type='yolov4' if 'yolov4' in cfg else 'yolov3')
            # focal-loss style bias prior on the preceding head conv (models.py:135-144)
            try:
                b = module_list[-1][0].bias.view(modules.na, -1)
                b.data[:, 4] += -4.5
                b.data[:, 5:] += math.log(0.6 / (modules.nc - 0.99))
                module_list[-1][0].bias = torch.nn.Parameter(b.view(-1), requires_grad=True)
            except Exception as e:  # noqa: BLE001 - the reference only warns here
                print('WARNING: smart bias initialization failure.', e)
        else:
            print("Warning: Unrecognized Layer Type: " + kind)

        module_list.append(modules)
        out_filters.append(filters)

    routs_binary = [False] * len(modules_defs)
    for i in routs:
        routs_binary[i] = True
    return module_list, routs_binary, net_infos


class YOLOLayer(nn.Module):
    """Detection head post-processing (reference models.py:158-258): permute to (bs, na, ny, nx, no) and,
    in eval mode, decode boxes with the yolov3 or yolov4 formulas.  Runs as one native kernel."""

    def __init__(self, anchors, nc, img_size, stride, bf_type='yolov3'):
        super().__init__()
        self.anchors = torch.Tensor(anchors)
        self.stride = stride
        self.na = len(anchors)
        self.nc = nc
        self.no = nc + 5
        self.nx, self.ny, self.ng = 0, 0, (0, 0)
        self.anchor_vec = self.anchors / self.stride
        self.anchor_wh = self.anchor_vec.view(1, self.na, 1, 1, 2)
        self.bf_type = bf_type
        self.grid = None

    def create_grids(self, ng=(13, 13), device="cpu"):
        """Keeps the reference's bookkeeping attributes current (nx, ny, ng, grid, anchor_vec device);
        the decode kernel derives the grid offsets from the thread index and does not read `grid`."""
        self.nx, self.ny = ng
        self.ng = torch.tensor(ng, dtype=torch.float)
        if not self.training:
            yv, xv = torch.meshgrid([torch.arange(self.ny, device=device), torch.arange(self.nx, device=device)],
                                    indexing='ij')
            self.grid = torch.stack((xv, yv), 2).view((1, 1, self.ny, self.nx, 2)).float()
        if self.anchor_vec.device != device:
            self.anchor_vec = self.anchor_vec.to(device)
            self.anchor_wh = self.anchor_wh.to(device)

    def forward(self, p):
        if self.bf_type not in ('yolov3', 'yolov4'):
            raise TypeError("bounding box predication error")
        _ops._require_cuda(p, "YOLOLayer.forward")
        bs, _, ny, nx = p.shape
        if (self.nx, self.ny) != (nx, ny) or self.grid is None:
            self.create_grids((nx, ny), p.device)
        head = p.detach().float().permute(0, 2, 3, 1).contiguous()  # layout only; arithmetic is native
        p_out = torch.empty((bs, self.na, ny, nx, self.no), dtype=torch.float32, device=p.device)
        rows = self.na * ny * nx
        io = None if self.training else torch.empty((bs, rows, self.no), dtype=torch.float32, device=p.device)
        _ops.yolo_decode(head, head.shape[-1], p_out, io, N=bs, ny=ny, nx=nx, na=self.na, no=self.no,
                         anchor_vec=self.anchor_vec.contiguous(), stride=self.stride,
                         v4=self.bf_type == 'yolov4', rows_total=rows, row_off=0)
        return p_out if self.training else (io, p_out)


class YOLO(nn.Module):
    """Darknet model built from a cfg file (reference models.py:261-315).

    ``forward(x, y=None)``: x = visible frames, y = LWIR frames, NCHW float (values in [0,1]) or uint8
    (raw pixels; the /255 of the reference's callers is then fused into the stem kernel) CUDA tensors.
    Training mode returns the list of raw head tensors (differentiable: forward with batch-statistics BatchNorm and
    the whole backward run as native kernels inside one autograd node, dyk/train_plan.py), eval mode
    ``(cat(io, 1), tuple(p))``.

    Extra knobs (not in the reference): ``compute_dtype`` (torch.float16 | torch.bfloat16; under
    ``torch.autocast`` the autocast dtype wins) and ``use_cuda_graph``.
    """

    def __init__(self, cfg, img_size=(416, 416), verbose=False):
        super().__init__()
        self.module_defs = parse_model_cfg(cfg)
        self.module_list, self.routs, self.net_info = create_modules(self.module_defs, img_size, cfg)
        self.yolo_layers = get_yolo_layers(self)
        self.cfg = cfg
        self.compute_dtype = torch.float16
        self.use_cuda_graph = True
        self._plans = _plan.PlanCache(self)
        self._train_plans = _train_plan.TrainPlanCache(self)
        self.info(verbose)

    def get_yolo_layers(self):
        return [i for i, module in enumerate(self.module_list) if module.__class__.__name__ == 'YOLOLayer']

    def info(self, verbose=False):
        torch_utils.model_info(self, verbose)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if '_plans' in self.__dict__:
            self._plans.invalidate()  # parameter storage may have moved
        if '_train_plans' in self.__dict__:
            self._train_plans.invalidate()
        return out

    def refresh_weights(self):
        """Re-pack the native copies of the parameters (packed 16-bit weights, folded BatchNorm, captured graphs) on the
        next forward.  In-place updates through the tensors themselves (optimizer.step, load_state_dict, copy_ under
        no_grad) are detected by their version counters; writes through ``param.data`` are NOT — call this after them."""
        if '_plans' in self.__dict__:
            self._plans.mark_stale()

    def forward(self, x, y=None, input_size=None):
        """input_size=(H, W) (not in the reference): the frames are bilinearly resized to H x W on the fly inside the stem
        convolutions — the fused form of the multi-scale training loop's `imgs = F.interpolate(imgs, size=ns,
        mode='bilinear', align_corners=False)` (train_utils/kaist_train_eval_utils.py:59-71); pass the ORIGINAL frames."""
        di = "second_index" in self.net_info and y is not None
        if self.training:
            if "second_index" in self.net_info and y is None:
                raise ValueError("this cfg defines second_index: call model(visible, lwir) with both modalities")
            return self._train_plans.run(x, y if di else None, input_size)   # list of raw head tensors with a grad_fn
        io, p = self._plans.run(x, y if di else None, input_size)
        return io, p


def load_darknet_weights(model, weights, cutoff=-1):
    """Loads a Darknet ``.weights`` blob (header int32x3 + int64, then per conv: BN bias/weight/mean/var or
    conv bias, then conv weights) into the model — reference models.py:318-364."""
    assert weights.endswith('.weights'), "weights file must end with '.weights'"
    with open(weights, 'rb') as f:
        model.version = np.fromfile(f, dtype=np.int32, count=3)
        model.seen = np.fromfile(f, dtype=np.int64, count=1)
        blob = np.fromfile(f, dtype=np.float32)

    ptr = 0

    def take(dst):
        nonlocal ptr
        n = dst.numel()
        if ptr + n > blob.size:
            raise ValueError(f"{weights}: file ends after {blob.size} floats, the model needs more")
        with torch.no_grad():   # an in-place copy on the tensor itself bumps its version counter (a `.data` write does not)
            dst.copy_(torch.from_numpy(blob[ptr:ptr + n]).view_as(dst))
        ptr += n

    for mdef, module in zip(model.module_defs[:cutoff], model.module_list[:cutoff]):
        if mdef['type'] != 'convolutional':
            continue
        conv = module[0]
        if mdef['batch_normalize']:
            bn = module[1]
            for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var):
                take(t)
        else:
            take(conv.bias)
        take(conv.weight)
    model.refresh_weights()
