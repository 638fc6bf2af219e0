# Shadow of reference allocate_cuda_device.py:6-7 (hard-coded cuda:1): device from the env.
import os
import torch as tch

def allocate_cuda():
    return tch.device(os.environ.get("EESEG_ORACLE_DEVICE", "cpu"))
