"""Host-side helpers models.py imports (subset of the reference's build_utils/torch_utils.py:18-74 that
the model API touches).  Pure plumbing — nothing here is on the compute path."""
import math
import time
from copy import deepcopy

import torch
import torch.nn as nn


def init_seeds(seed=0):
    torch.manual_seed(seed)


def time_synchronized():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def select_device(device='cuda:0'):
    """Single GPU (or 'cpu' for host-side tooling; the model itself only runs on a B200)."""
    cpu_request = device.lower() == 'cpu'
    if device and not cpu_request:
        assert torch.cuda.is_available(), 'CUDA unavailable, invalid device {} requested'.format(device)
    dev = torch.device(device)
    if not cpu_request and torch.cuda.is_available():
        di = 0 if dev.index is None else dev.index
        dp = torch.cuda.get_device_properties(di)
        print("Using torch %s CUDA:%d (%s, %dMB)" % (torch.__version__, di, dp.name, dp.total_memory / 1024 ** 2))
    else:
        print(f'Using torch {torch.__version__} CPU')
    return dev


def model_info(model, verbose=False):
    """One-line (or per-parameter) summary, as reference torch_utils.py:55-74 minus the optional thop pass."""
    n_p = sum(x.numel() for x in model.parameters())
    n_g = sum(x.numel() for x in model.parameters() if x.requires_grad)
    if verbose:
        print('%5s %40s %9s %12s %20s %10s %10s' % ('layer', 'name', 'gradient', 'parameters', 'shape', 'mu', 'sigma'))
        for i, (name, p) in enumerate(model.named_parameters()):
            name = name.replace('module_list.', '')
            print('%5g %40s %9s %12g %20s %10.3g %10.3g' %
                  (i, name, p.requires_grad, p.numel(), list(p.shape), p.mean(), p.std()))
    print(f"Model Summary: {len(list(model.modules()))} layers, {n_p} parameters, {n_g} gradients")


def initialize_weights(model):
    """BatchNorm eps / momentum and in-place activations as the reference sets them (torch_utils.py:23-32); the native
    kernels read `bn.eps` / `bn.momentum` from the modules, so the change takes effect on the next forward."""
    for m in model.modules():
        if type(m) is nn.BatchNorm2d:
            m.eps, m.momentum = 1e-4, 0.03
        elif type(m) in (nn.LeakyReLU, nn.ReLU, nn.ReLU6):
            m.inplace = True


class ModelEMA:
    """Exponential moving average of a model's state_dict with the reference's warm-up ramp d(n) = decay*(1 - e^(-n/2000))
    (torch_utils.py:77-126).  Host-side bookkeeping between optimizer steps, not part of the forward/backward path: the
    average lives in a deep copy of the model (eval mode), whose native plans re-pack themselves when its tensors change."""

    def __init__(self, model, decay=0.9999, device=''):
        self.ema = deepcopy(model).eval()
        self.updates = 0
        self.decay = lambda n: decay * (1 - math.exp(-n / 2000))
        self.device = device
        if device:
            self.ema.to(device=device)
        for p in self.ema.parameters():
            p.requires_grad_(False)

    @staticmethod
    def _unwrap(m):
        return m.module if isinstance(m, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)) else m

    def update(self, model):
        self.updates += 1
        d = self.decay(self.updates)
        with torch.no_grad():
            src = self._unwrap(model).state_dict()
            avg = self._unwrap(self.ema).state_dict()
            keys = [k for k, v in avg.items() if v.dtype.is_floating_point]
            dst = [avg[k] for k in keys]
            torch._foreach_mul_(dst, d)
            torch._foreach_add_(dst, [src[k].detach().to(avg[k].device) for k in keys], alpha=1.0 - d)

    def update_attr(self, model):
        for k, v in model.__dict__.items():
            if not k.startswith('_'):
                setattr(self.ema, k, v)
