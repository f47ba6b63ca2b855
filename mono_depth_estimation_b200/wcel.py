"""WCEL_Loss (reference criteria.py:839-863), the other half of ModelLoss.

SURVEY 8(f) rank 1 ("next" row): not part of the hot path named by BASELINE.json. Until its
channel-streaming kernel lands, constructing it states that plainly instead of silently running
an unfused PyTorch chain.
"""
from __future__ import annotations

import torch.nn as nn


class WCEL_Loss(nn.Module):
    def __init__(self, args):
        super().__init__()
        raise NotImplementedError(
            "WCEL_Loss (reference criteria.py:839-863) is a 'next' row of the scope table (SURVEY 8f); "
            "use criteria.VNL_Loss directly for the virtual-normal term")
