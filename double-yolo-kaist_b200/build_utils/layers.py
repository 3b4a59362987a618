"""Op modules of the Darknet graph — same class names, constructor signatures, parameter names and
``state_dict`` keys as the reference's build_utils/layers.py, so checkpoints and the class-name
dispatch in ``YOLO.forward`` (reference models.py:292-296) keep working.

Inside ``models.YOLO`` these objects are *parameter containers*: the execution plan (dyk/plan.py)
reads their weights and launches the fused sm_100a kernels.  Called on their own (``module(x)`` /
``module(x, outputs)`` with NCHW CUDA tensors, as reference code may do) they run the same native
kernels through ``dyk.ops``; there is no PyTorch-eager arithmetic path in this file.
"""
import math

import torch
import torch.nn as nn

from dyk import ops as _ops


def make_divisible(v, divisor):
    """Round up to a multiple of `divisor` (reference layers.py:9-11)."""
    return math.ceil(v / divisor) * divisor


class FeatureConcat(nn.Module):
    """[route]: channel concat of earlier outputs; a single index returns that tensor itself
    (an alias, reference layers.py:32-44)."""

    def __init__(self, layers):
        super().__init__()
        self.layers = layers
        self.multiple = len(layers) > 1

    def forward(self, x, outputs):
        if not self.multiple:
            return outputs[self.layers[0]]
        return _ops.concat_channels([outputs[i] for i in self.layers])


class WeightedFeatureFusion(nn.Module):
    """[shortcut]: x + sum(outputs[l]) or, with ``weights_type``, the sigmoid-weighted sum
    ``w = sigmoid(self.w) * 2/n`` (reference layers.py:47-85)."""

    def __init__(self, layers, weight=False):
        super().__init__()
        self.layers = layers
        self.weight = weight
        self.n = len(layers) + 1
        if weight:
            self.w = nn.Parameter(torch.zeros(self.n), requires_grad=True)

    def forward(self, x, outputs):
        return _ops.weighted_fusion(x, [outputs[l] for l in self.layers], self.w if self.weight else None)


def _activation_module(name):
    table = {
        'mish': lambda: nn.Mish(inplace=True), 'relu': lambda: nn.ReLU(inplace=True),
        'relu6': lambda: nn.ReLU6(inplace=True), 'leaky': lambda: nn.LeakyReLU(0.1, inplace=True),
        'hard-sigmoid': lambda: nn.Hardsigmoid(inplace=True), 'hard-swish': lambda: nn.Hardswish(inplace=True),
    }
    return table[name]() if name in table else None


def activation_name(module) -> str:
    """Inverse of _activation_module for the plan compiler."""
    names = {nn.Mish: 'mish', nn.ReLU: 'relu', nn.ReLU6: 'relu6', nn.LeakyReLU: 'leaky',
             nn.Hardsigmoid: 'hard-sigmoid', nn.Hardswish: 'hard-swish'}
    return names.get(type(module), 'linear')


class ConvBnActivation(nn.Module):
    """Conv2d (+BatchNorm2d) (+activation) held in ``self.conv`` (ModuleList), as reference
    layers.py:88-122 — used by Inception."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, pad=0, groups=1, activation="leaky",
                 bn=True):
        super().__init__()
        self.conv = nn.ModuleList()
        self.conv.append(nn.Conv2d(in_channels, out_channels, kernel_size, stride,
                                   padding=kernel_size // 2 if pad else 0, groups=groups, bias=not bn))
        if bn:
            self.conv.append(nn.BatchNorm2d(out_channels))
        act = _activation_module(activation)
        if act is not None:
            self.conv.append(act)
        elif activation != 'linear':
            print("activate error: unknown activation {}".format(activation))

    def forward(self, x):
        conv = self.conv[0]
        bn = self.conv[1] if len(self.conv) > 1 and isinstance(self.conv[1], nn.BatchNorm2d) else None
        return _ops.conv_bn_act(x, conv, bn, activation_name(self.conv[-1]))


class Inception(nn.Module):
    """Four-branch Inception block (reference layers.py:148-172): 1x1 | 1x1-3x3 | 1x1-3x3-3x3 |
    maxpool3-1x1, concatenated."""

    def __init__(self, in_channels, n1x1, n3x3_reduce, n3x3, n5x5_reduce, n5x5, pool_proj):
        super().__init__()
        self.branch1 = nn.Sequential(ConvBnActivation(in_channels, n1x1, kernel_size=1))
        self.branch2 = nn.Sequential(ConvBnActivation(in_channels, n3x3_reduce, kernel_size=1),
                                     ConvBnActivation(n3x3_reduce, n3x3, kernel_size=3, pad=1))
        self.branch3 = nn.Sequential(ConvBnActivation(in_channels, n5x5_reduce, kernel_size=1),
                                     ConvBnActivation(n5x5_reduce, n5x5, kernel_size=3, pad=1),
                                     ConvBnActivation(n5x5, n5x5, kernel_size=3, pad=1))
        self.branch4 = nn.Sequential(nn.MaxPool2d(kernel_size=3, stride=1, padding=1),
                                     ConvBnActivation(in_channels, pool_proj, kernel_size=1))

    def forward(self, x):
        b4 = self.branch4[1](_ops.maxpool(x, 3, 1))
        return _ops.concat_channels([self.branch1(x), self.branch2(x), self.branch3(x), b4])


class SqueezeExcitation(nn.Module):
    """Channel attention: x * hardsigmoid(fc2(relu(fc1(mean_hw(x))))) (reference layers.py:175-190)."""

    def __init__(self, in_channels: int, squeeze_factor: int = 4):
        super().__init__()
        squeeze_channel = make_divisible(in_channels // squeeze_factor, 8)
        self.fc1 = nn.Conv2d(in_channels, squeeze_channel, 1)
        self.fc2 = nn.Conv2d(squeeze_channel, in_channels, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _ops.squeeze_excitation(x, self.fc1, self.fc2)


class DepthwiseSeparableConv2d(nn.Module):
    """depthwise kxk (padding fixed to 1) -> BN -> ReLU6 -> pointwise 1x1 -> BN -> ReLU6, children in
    ``self.conv`` = Sequential indices 0,1,3,4 (reference layers.py:218-234)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, kernel_size, stride, 1, groups=in_channels, bias=False),
            nn.BatchNorm2d(in_channels),
            nn.ReLU6(inplace=True),
            nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU6(inplace=True))

    def forward(self, x):
        x = _ops.conv_bn_act(x, self.conv[0], self.conv[1], 'relu6')
        return _ops.conv_bn_act(x, self.conv[3], self.conv[4], 'relu6')


# ---- import compatibility for names the reference file also defines but create_modules never uses
class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        if self.d != 1:
            raise NotImplementedError("Concat: only channel concatenation is on the hot path")
        return _ops.concat_channels(list(x))


class ResBlock(nn.Module):
    """`block_nums` x (1x1 conv -> 3x3 conv [+ input]) — defined by the reference (layers.py:125-145) but used by no cfg;
    kept for `from build_utils.layers import *`.  Each conv is a native ConvBnActivation, the add a native fused add."""

    def __init__(self, in_channels, filter_1, out_channels, block_nums=1, activation='mish', shortcut=True):
        super().__init__()
        self.shortcut = shortcut
        self.module_list = nn.ModuleList()
        for _ in range(block_nums):
            # the reference passes (…, 1, 1, 1, activation, True) positionally: kernel 1, stride 1, pad 1, groups=activation —
            # i.e. it cannot be constructed there; the evident intent is built here
            self.module_list.append(nn.ModuleList([
                ConvBnActivation(in_channels, filter_1, 1, 1, 1, activation=activation, bn=True),
                ConvBnActivation(filter_1, out_channels, 3, 1, 1, activation=activation, bn=True)]))

    def forward(self, x):
        for pair in self.module_list:
            y = x
            for conv in pair:
                y = conv(y)
            x = _ops.weighted_fusion(y, [x], None) if self.shortcut else y
        return x


class SEInceptionFusion(nn.Module):
    """concat(layers) -> 1x1 conv -> [Inception] -> [SqueezeExcitation] (reference layers.py:193-215; no cfg block type
    builds it).  Composed of the native modules above."""

    def __init__(self, in_channels, out_channels, layers, inception=False, icp_param_list=(), tmse=False, squeeze_factor=4):
        super().__init__()
        self.concat = FeatureConcat(layers)
        self.enhance = nn.ModuleList([ConvBnActivation(in_channels, out_channels, kernel_size=1)])
        if inception:
            self.enhance.append(Inception(out_channels, *icp_param_list))
        if tmse:
            self.enhance.append(SqueezeExcitation(out_channels, squeeze_factor))

    def forward(self, x, outputs):
        y = self.concat(x, outputs)
        for m in self.enhance:
            y = m(y)
        return y


class MixConv2d(nn.Module):
    """Parallel convolutions with different kernel sizes, concatenated (reference layers.py:237-268; unused by every cfg).
    Channel split rules as in the reference: 'equal_ch' or 'equal_params'."""

    def __init__(self, in_ch, out_ch, k=(3, 5, 7), stride=1, dilation=1, bias=True, method='equal_params'):
        super().__init__()
        import numpy as np
        groups = len(k)
        if method == 'equal_ch':
            idx = torch.linspace(0, groups - 1E-6, out_ch).floor()
            ch = [int((idx == g).sum()) for g in range(groups)]
        else:
            a = np.eye(groups + 1, groups, k=-1)
            a -= np.roll(a, 1, axis=1)
            a *= np.array(k) ** 2
            a[0] = 1
            ch = np.linalg.lstsq(a, [out_ch] + [0] * groups, rcond=None)[0].round().astype(int)
        if dilation != 1:
            raise _ops.nat.NativeError("MixConv2d: dilated convolutions have no native kernel")
        self.m = nn.ModuleList([nn.Conv2d(in_ch, int(ch[g]), k[g], stride, k[g] // 2, bias=bias) for g in range(groups)])

    def forward(self, x):
        return _ops.concat_channels([_ops.conv_bn_act(x, m, None, 'linear') for m in self.m])


class _NativeActivation(nn.Module):
    """Stand-alone activation modules the reference defines next to nn.Mish / nn.Hardswish (layers.py:271-320); the
    model never instantiates them (create_modules uses the torch.nn classes as parameter-free markers)."""
    act = 'linear'

    def forward(self, x):
        return _ops.activation(x, self.act)


class Mish(_NativeActivation):
    act = 'mish'


class MemoryEfficientMish(Mish):
    pass


class HardSwish(_NativeActivation):
    act = 'hard-swish'


class Swish(nn.Module):
    def forward(self, x):
        raise _ops.nat.NativeError("Swish (x*sigmoid(x)) is not an activation of any Darknet cfg block and has no native kernel")


class MemoryEfficientSwish(Swish):
    pass
