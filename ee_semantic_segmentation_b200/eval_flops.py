"""Drop-in for the analysis helpers of the reference's eval_flops.py: `check_flops` (:15-26) and
`count_flops` (:28-50). The reference traces with third-party `pthflops` on the real device per
call; here FLOPs are counted once on meta tensors with torch.utils.flop_counter (no device work)."""
import copy

import torch as tch
from torch import nn
from torch.utils.flop_counter import FlopCounterMode


def check_flops(aux_model, img_dim, channels, device=None):
    if isinstance(aux_model, list):
        aux_model = nn.Sequential(*aux_model)
    m = copy.deepcopy(aux_model).to('meta').eval()
    if isinstance(img_dim, (list, tuple, tch.Size)):
        x = tch.empty(1, channels, img_dim[0], img_dim[1], device='meta')
    else:
        x = tch.empty(1, channels, img_dim, img_dim, device='meta')
    with tch.no_grad(), FlopCounterMode(display=False) as fc:
        m(x)
    return fc.get_total_flops()


def count_flops(net, device=None, img_dim=256, channels=3):
    x_dim, y_dim = (img_dim, img_dim) if isinstance(img_dim, int) else img_dim
    main_flops, branch_flops = [], []
    X = tch.empty(1, channels, x_dim, y_dim, device='meta')
    for i in range(net.n_branches + 1):
        sec = copy.deepcopy(net.base_model[i]).to('meta').eval()
        main_flops.append(check_flops(net.base_model[i], X.shape[-2:], X.shape[1]))
        with tch.no_grad():
            X = sec(X)
        head = net.branches[i] if i < net.n_branches else net.classifier
        branch_flops.append(check_flops(head, X.shape[-2:], X.shape[1]))
    for i in range(1, len(main_flops)):
        main_flops[i] += main_flops[i - 1]
    return [i + j for i, j in zip(main_flops, branch_flops)]
