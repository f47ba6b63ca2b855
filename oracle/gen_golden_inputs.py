"""ORACLE (test infrastructure): seeded synthetic inputs shared by oracle/gen_golden.py and the GPU parity tests
(no reference imports here: this file travels to the GPU box)."""
import torch


def stdepth_inputs(seed, B, C, H, W):
    """Layered-depth batch: 8 (or 16) colour/alpha channels in [0,1], depth channels in (0, 1] with holes, an alpha
    plane with holes; pred = targ + noise, positive on the depth channels."""
    g = torch.Generator().manual_seed(seed)
    targ = torch.rand((B, C, H, W), generator=g)
    d0 = 8 if C == 10 else 16
    targ[:, d0:] = targ[:, d0:] * 0.95 + 0.05
    targ[:, d0:][torch.rand((B, C - d0, H, W), generator=g) < 0.3] = 0.0
    targ[:, d0][torch.rand((B, H, W), generator=g) < 0.05] = 0.005          # in maskD, outside silog's own mask (> 1e-2)
    rgba = torch.rand((B, 4, H, W), generator=g)
    rgba[:, 3][torch.rand((B, H, W), generator=g) < 0.35] = 0.0
    pred = targ + torch.randn((B, C, H, W), generator=g) * 0.1
    pred[:, d0:] = pred[:, d0:].abs() + 0.02
    pred[0, 0, 2, 3:6] = targ[0, 0, 2, 3:6]                                 # exact ties: sign(0) = 0
    pred[0, :3, 5, 5] = 0.0                                                 # zero front vector: norm gradient 0
    return pred, targ, rgba
