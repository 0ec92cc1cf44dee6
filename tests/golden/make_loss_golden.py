"""Generates tests/golden/loss.npz with the REFERENCE's own utils/loss_utils.py (l1_loss, ssim) on CPU: loss value
and autograd gradient of  (1 - 0.2) * l1 + 0.2 * (1 - ssim)  (train.py:192-196) for two seeded image pairs, one
with sizes that are not multiples of the kernel tile.  Run in the authoring container."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_decode_golden import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    import_reference()
    from utils.loss_utils import l1_loss, ssim
    g = torch.Generator().manual_seed(99)
    out = {}
    for n, (H, W) in enumerate([(48, 64), (37, 53)]):
        gt = torch.rand(3, H, W, generator=g)
        img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1).requires_grad_()
        l1 = l1_loss(img, gt)
        ss = ssim(img, gt)
        loss = (1.0 - 0.2) * l1 + 0.2 * (1.0 - ss)
        loss.backward()
        out.update({f"p{n}.img": img.detach().numpy(), f"p{n}.gt": gt.numpy(), f"p{n}.loss": loss.item(), f"p{n}.l1": l1.item(),
                    f"p{n}.ssim": ss.item(), f"p{n}.grad": img.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)
    print({k: v for k, v in out.items() if not hasattr(v, "shape")})


if __name__ == "__main__":
    main()
