import torch.nn.functional as F  # noqa: F401  (block.py:3 imports F from here)
from oracle.flux_oracle import Attention  # noqa: F401
