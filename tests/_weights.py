"""Deterministic, machine-independent weights for parity tests.

Every tensor of a module's state_dict is filled from a torch CPU generator seeded by
crc32(key name), so the reference (in the build container), the oracle and the CUDA product
(on the GPU box) can all be given identical weights without shipping checkpoints.
"""
import math
import zlib

import torch


@torch.no_grad()
def fill_deterministic(module, salt=""):
    sd = module.state_dict()
    for key, t in sd.items():
        g = torch.Generator().manual_seed(zlib.crc32((salt + key).encode()))
        if key.endswith("num_batches_tracked"):
            t.zero_()
        elif key.endswith("running_mean"):
            t.zero_()
        elif key.endswith("running_var"):
            t.fill_(1.0)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            t.copy_(torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in))
        elif key.endswith("weight"):          # norm scale
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
        else:                                  # biases
            t.copy_(0.05 * torch.randn(t.shape, generator=g))
    module.load_state_dict(sd)
    return module
