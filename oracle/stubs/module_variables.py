# Shadow of the reference's un-importable `module_variables` registry.
from torch import nn
