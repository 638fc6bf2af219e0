# Stub for the reference's missing `common_header` module.
# TEST INFRASTRUCTURE ONLY: lets /root/reference import in the build container.
import os, re, sys
import numpy as np
import torch as tch
from pandas import DataFrame
from typing import List, Callable, Optional, Tuple
