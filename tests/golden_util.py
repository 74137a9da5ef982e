"""Loader for the committed reference fixtures (tests/golden/*.npz, written by oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.cfg = json.loads(str(z["cfg"]))
        self.inp, self.sd, self.out, self.grad_in, self.grad_sd = {}, {}, {}, {}, {}
        for k in z.files:
            if k == "cfg":
                continue
            t = torch.from_numpy(z[k])
            for pre, dst in (("in.", self.inp), ("sd.", self.sd), ("out.", self.out), ("grad.in.", self.grad_in),
                             ("grad.sd.", self.grad_sd)):
                if k.startswith(pre):
                    dst[k[len(pre):]] = t
                    break

    def sd_as(self, dtype=None, device=None):
        out = {}
        for k, v in self.sd.items():
            if v.is_floating_point() and dtype is not None:
                v = v.to(dtype)
            out[k] = v.to(device) if device is not None else v
        return out


def names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))


def rel_err(a, b):
    """global relative L2 error of a against the reference b"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
