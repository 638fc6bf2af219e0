"""Drop-in for the reference's from_deepv3.py (v1 of the branchy model: fixed DeepLabHead
branches, 21 classes — from_deepv3.py:30-125). Same class as from_deepv3_new, without
`branch_params`; `ee_dnn_op*.py` import the model from this module name."""
from .from_deepv3_new import branchyDeepv3 as _branchyDeepv3
from .from_deepv3_new import get_base_model  # noqa: F401


class branchyDeepv3(_branchyDeepv3):
    def __init__(self, base_name, base_type, n, img_dim, count_branches=True, skip=0, **kw):
        super().__init__(base_name, base_type, n, img_dim, count_branches, skip, None, **kw)
