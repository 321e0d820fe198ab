from typing import Optional

import torch

Adj = torch.Tensor
OptTensor = Optional[torch.Tensor]
