"""Host-side frame augmentations used by the VID (continuity) loss: `warp` (dalle_bert.py:204-238) builds a negative
clip per sample by one of four corruptions.  Data-side torch ops (not hot-path kernels); RNG sources follow the
reference (numpy for the strategy / partner choice, python `random` for frame indices, torch for colour / affine)."""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F


def _non_identity_perm(n):
    """A permutation of range(n) that is not the identity (dalle_bert.py:93-108)."""
    if n < 6:
        from itertools import permutations
        return list(random.choice(list(permutations(range(n)))[1:]))
    ident = torch.arange(n)
    while True:
        perm = torch.randperm(n)
        if (perm != ident).any():
            return perm


def _colour_shift(frame):
    shift = (torch.rand(1) - 0.5).to(frame.device)
    which = random.randint(0, 3)
    delta = torch.zeros_like(frame)
    if which == 0:
        delta += shift
    else:
        delta[which - 1] += shift
    return torch.clamp(frame + delta, 0, 1)


def _affine(frame, angle_deg=30, trans=0.1, scale=0.1):
    ang = math.pi * angle_deg / 180.0
    a = torch.empty(1).uniform_(-ang, ang)
    tx, ty = torch.empty(1).uniform_(-trans, trans), torch.empty(1).uniform_(-trans, trans)
    sc = torch.empty(1).uniform_(1.0 - scale, 1.0 + scale)
    theta = torch.tensor([[sc * torch.cos(a), sc * torch.sin(-a), tx], [sc * torch.sin(a), sc * torch.cos(a), ty]]).unsqueeze(0)
    x = frame.unsqueeze(0)
    grid = F.affine_grid(theta, x.size(), align_corners=False).to(x.device)
    return F.grid_sample(x, grid, padding_mode="reflection", align_corners=False)[0]


def warp_video_with_color(video):
    """dalle_bert.py:140-158: one colour shift per clip (all channels or one of R / G / B), the same for every frame of
    the clip.  video [n, t, 3, h, w] in [0, 1].  RNG order per clip as in the reference: torch.rand(1) (CPU generator), then
    python random.randint(0, 3)."""
    out = []
    for n in range(video.shape[0]):
        x = video[n]
        shift = (torch.rand(1) - 0.5).to(x.device)
        which = random.randint(0, 3)
        delta = torch.zeros_like(x)
        if which == 0:
            delta += shift
        else:
            delta[:, which - 1] += shift
        out.append(torch.clamp(x + delta, 0, 1))
    return torch.stack(out)


def augment_visual(visual, visual_aug_mode):
    """BERT.forward's training-time visual-control augmentation (dalle_bert.py:940-944): with probability 0.9 every
    control frame but the first of each sample gets the clip's colour shift ('motion_color')."""
    # any other mode is silently ignored by the reference, and draws nothing (`==` short-circuits the `and`)
    if visual_aug_mode == "motion_color" and random.random() < 0.9:
        out = visual.detach().clone()
        out[:, 1:] = warp_video_with_color(visual[:, 1:])
        return out
    return visual


def warp(x, vid_strategy_prob=(0.25, 0.25, 0.25, 0.25)):
    """x [b,t,c,h,w] -> corrupted copy: 0 frame from another clip, 1 shuffled frames, 2 colour shift, 3 affine warp."""
    b, t = x.shape[:2]
    out = []
    for i in range(b):
        strategy = int(np.random.choice(range(4), p=vid_strategy_prob))
        y = x[i].detach().clone()
        if strategy == 0:
            other = int(np.random.choice(list(set(range(b)) - {i})))
            j1, j2 = random.randint(0, t - 1), random.randint(0, t - 1)
            y[j1] = x[other, j2]
        elif strategy == 1:
            y = x[i, _non_identity_perm(t)].detach().clone()
        elif strategy == 2:
            j1 = random.randint(0, t - 1)
            y[j1] = _colour_shift(y[j1])
        else:
            j1 = random.randint(0, t - 1)
            y[j1] = _affine(y[j1])
        out.append(y)
    return torch.stack(out, 0)
