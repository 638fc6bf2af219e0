# Stand-in for third-party `pthflops.count_ops` (not installed, unpinned in the reference).
# Counts with torch's FlopCounterMode; used only to place branches the way the survey probe did.
# The model is traced in eval mode (train-mode BatchNorm rejects the 1x1 pooled ASPP tensor at
# batch 1) and its mode is restored afterwards.
import torch
from torch.utils.flop_counter import FlopCounterMode

def count_ops(model, x, print_readable=False, verbose=False, **kw):
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad(), FlopCounterMode(display=False) as fc:
            model(x)
    finally:
        model.train(was_training)
    return fc.get_total_flops(), None
