"""Shared parity helpers.  Tolerances are stated per tensor as ||a-b||/||b|| against the fp32 reference and are
calibrated by the oracle's own bf16-operand emulation of the same computation (operands rounded to bf16, fp32
accumulation): a tensor passes when  err(ours) <= FLOOR + FACTOR * err(emulation)."""
import os

import numpy as np
import torch

from oracle import savqa_oracle as O

FLOOR = 3e-3   # bf16 has 8 mantissa bits: 2^-9 = 2e-3 per rounding; outputs that see one rounding sit at ~1e-3
FACTOR = 2.5


def t(a):
    return torch.from_numpy(np.asarray(a))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def check(name, got, ref, emu=None, floor=FLOOR, factor=FACTOR, mask=None):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    if mask is not None:
        got, ref = got.clone(), ref.clone()
        got[mask] = 0
        ref[mask] = 0
    err = O.rel_err(got, ref)
    tol = floor
    if emu is not None:
        emu = emu.detach().float().cpu()
        if mask is not None:
            emu = emu.clone()
            emu[mask] = 0
        tol = floor + factor * O.rel_err(emu, ref)
    assert err <= tol, f"{name}: rel err {err:.3e} > tol {tol:.3e}"
    return err


def grads_of(module, keys):
    named = dict(module.named_parameters())
    return {k: (named[k].grad.detach().cpu() if named[k].grad is not None else torch.zeros_like(named[k]).cpu()) for k in keys}


def set_params(module, P):
    sd = module.state_dict()
    diff = set(sd) ^ set(P)
    assert not diff, sorted(diff)[:5]
    module.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
