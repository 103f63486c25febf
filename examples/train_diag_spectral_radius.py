#!/usr/bin/env python
"""dp_playground in miniature: learn one diagonal Q_delta (the reference's `Params` "model", dp_playground.py:478-560)
by gradient descent on the mean spectral radius over freshly sampled lambdas - forward and backward both run the
hand-written kernels (`SpectralRadiusLoss.differentiable`).

    python examples/train_diag_spectral_radius.py --M 5 --steps 300 --batch 65536
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdc_gym_b200.loss import SpectralRadiusLoss  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=5)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(args.seed)
    loss_fn = SpectralRadiusLoss(args.M, 1.0, "diag")
    theta = torch.full((args.M,), 0.5, dtype=torch.float64, device=dev, requires_grad=True)
    opt = torch.optim.Adam([theta], lr=args.lr)
    history = []
    for step in range(args.steps):
        lam = torch.complex(torch.rand(args.batch, dtype=torch.float64, device=dev, generator=gen) * -100.0,
                            torch.zeros(args.batch, dtype=torch.float64, device=dev))
        opt.zero_grad()
        loss = loss_fn.differentiable(lam, theta.expand(args.batch, args.M))
        loss.backward()
        opt.step()
        history.append(float(loss.detach()))
        if step % 50 == 0 or step == args.steps - 1:
            print(f"step {step:4d}  mean rho {history[-1]:.6f}  diag {theta.detach().cpu().numpy().round(4)}")
    return history, theta.detach().cpu().numpy()


if __name__ == "__main__":
    main()
