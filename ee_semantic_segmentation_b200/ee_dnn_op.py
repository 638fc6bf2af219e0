"""Drop-in for the operator class of the reference's ee_dnn_op.py (:40-118): the SIMILARITY-gated
twin of the entropy operator — an image leaves at exit i when metric(previous exit's argmax map,
this exit's argmax map) crosses the threshold. The skimage similarity metrics themselves
(sim_metrics.py) are CPU analysis code and out of scope: `metric` is any
callable taking two int64 [H,W] CPU tensors. The reference's `sel.threshold` typo (:84) is fixed. Result keys as in the
reference, including the `*_flops_2` variants that leave out the first branch (the reference image, :86-110)."""
import torch as tch

from . import ops
from .ee_dnn_op_ne import eval_ee_deeplabv3 as _EntropyOp
from .ee_dnn_op_ne import mIoU  # noqa: F401


class eval_ee_deeplabv3(_EntropyOp):
    def __call__(self, X):
        output = dict()
        inp_shape = X.shape[-2:]
        main_all, head_all = self._flop_table(X.shape)
        main_flops, branch_flops = [], []
        left, prev = False, None
        model = self.model
        if not X.is_cuda:
            raise RuntimeError('eval_ee_deeplabv3 needs a CUDA input (no CPU fallback)')
        X = X.unsqueeze(0)
        with tch.no_grad():
            for i in range(self.n):
                main_flops.append(main_all[i])
                X = model.run_section(i, X)
                if i not in self.ignore and not left:
                    low = model._plan(i).run(X)
                    branch_flops.append(head_all[i])
                    cur = ops.exit_gate(low, inp_shape, layout='NHWC', n_classes=model.num_classes,
                                        want_score=False).amax.squeeze(0).to(tch.int64).cpu()
                    if prev is not None:
                        t = self.metric(prev, cur)
                        if (t < self.threshold) if self.less_than else (t > self.threshold):
                            output['exit'] = cur
                            output['exit_flops'] = sum(branch_flops) + sum(main_flops)
                            output['exit_flops_2'] = sum(branch_flops[1:]) + sum(main_flops)
                            output['edge_flops'] = output['exit_flops']
                            output['edge_flops_2'] = output['exit_flops_2']
                            output['n'] = i + 1
                            left = True
                    prev = cur
                if not left and i == self.last_br:
                    output['edge_flops'] = sum(branch_flops) + sum(main_flops)
                    output['edge_flops_2'] = sum(branch_flops[1:]) + sum(main_flops)
            main_flops.append(main_all[self.n])
            X = model.run_section(self.n, X)
            main_flops.append(head_all[self.n])
            low = model._plan(self.n).run(X)
            Y = ops.exit_gate(low, inp_shape, layout='NHWC', n_classes=model.num_classes,
                              want_score=False).amax.squeeze(0).to(tch.int64).cpu()
        output['last'] = Y
        output['last_flops'] = sum(branch_flops) + sum(main_flops)
        output['last_flops_2'] = sum(branch_flops[1:]) + sum(main_flops)
        if not left:
            output['exit'] = Y
            output['exit_flops'] = output['last_flops']
            output['exit_flops_2'] = output['last_flops_2']
            output['n'] = self.n + 1
        return output
