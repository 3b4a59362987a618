"""Host-side helpers models.py imports (subset of the reference's build_utils/torch_utils.py:18-74 that
the model API touches).  Pure plumbing — nothing here is on the compute path."""
import time

import torch


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
