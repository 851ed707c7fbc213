"""DCGAN with two models, two fused optimizers and three backward passes per iteration under one GradScaler
(reference examples/dcgan/main_amp.py: the multi-model / multi-loss mixed-precision recipe). Synthetic 64x64 images.

    python main_amp.py --iters 200
"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from apex_b200.optimizers import FusedAdam  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--iters", type=int, default=100)
p.add_argument("--batch", type=int, default=64)
p.add_argument("--nz", type=int, default=100)
args = p.parse_args()
dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # CPU: the same recipe on the PyTorch reference paths
amp_dtype = torch.float16 if dev.type == "cuda" else torch.bfloat16


def G(nz, ngf=64):
    return nn.Sequential(nn.ConvTranspose2d(nz, ngf * 8, 4, 1, 0, bias=False), nn.BatchNorm2d(ngf * 8), nn.ReLU(True),
                         nn.ConvTranspose2d(ngf * 8, ngf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf * 4), nn.ReLU(True),
                         nn.ConvTranspose2d(ngf * 4, ngf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf * 2), nn.ReLU(True),
                         nn.ConvTranspose2d(ngf * 2, ngf, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf), nn.ReLU(True),
                         nn.ConvTranspose2d(ngf, 3, 4, 2, 1, bias=False), nn.Tanh())


def D(ndf=64):
    return nn.Sequential(nn.Conv2d(3, ndf, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 2), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 4), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 8), nn.LeakyReLU(0.2, True),
                         nn.Conv2d(ndf * 8, 1, 4, 1, 0, bias=False))


netG, netD = G(args.nz).to(dev), D().to(dev)
optD = FusedAdam(netD.parameters(), lr=2e-4, betas=(0.5, 0.999))
optG = FusedAdam(netG.parameters(), lr=2e-4, betas=(0.5, 0.999))
scaler = torch.amp.GradScaler(dev.type)
bce = nn.BCEWithLogitsLoss()
real = torch.randn(args.batch, 3, 64, 64, device=dev).tanh()
for it in range(args.iters):
    noise = torch.randn(args.batch, args.nz, 1, 1, device=dev)
    ones, zeros = torch.ones(args.batch, device=dev), torch.zeros(args.batch, device=dev)
    # (1) D on real, (2) D on fake
    optD.zero_grad()
    with torch.autocast(dev.type, dtype=amp_dtype):
        errD_real = bce(netD(real).view(-1).float(), ones)
        fake = netG(noise)
        errD_fake = bce(netD(fake.detach()).view(-1).float(), zeros)
    scaler.scale(errD_real).backward()
    scaler.scale(errD_fake).backward()
    scaler.step(optD)
    # (3) G
    optG.zero_grad()
    with torch.autocast(dev.type, dtype=amp_dtype):
        errG = bce(netD(fake).view(-1).float(), ones)
    scaler.scale(errG).backward()
    scaler.step(optG)
    scaler.update()
    if it % 20 == 0:
        print(f"[{it}/{args.iters}] Loss_D {(errD_real + errD_fake).item():.4f}  Loss_G {errG.item():.4f}  scale {scaler.get_scale():.0f}")
