"""PNA degree scalers with the reference's names and call signature (reference modules/pna/scalers.py:10-40):
`SCALERS[name](src, deg, avg_deg)`.  In the model identity / amplification / attenuation are folded into the
single-pass kernel gt_pna_reduce_* (csrc/pna.cu); these callables serve callers of the reference's module surface."""
import torch


def scale_identity(src, deg, avg_deg):
    return src


def scale_amplification(src, deg, avg_deg):
    return src * (torch.log(deg + 1) / avg_deg["log"])


def scale_attenuation(src, deg, avg_deg):
    scale = avg_deg["log"] / torch.log(deg + 1)
    scale = torch.where(deg == 0, torch.ones_like(scale), scale)
    return src * scale


def scale_linear(src, deg, avg_deg):
    return src * (deg / avg_deg["lin"])


def scale_inverse_linear(src, deg, avg_deg):
    scale = avg_deg["lin"] / deg
    scale = torch.where(deg == 0, torch.ones_like(scale), scale)
    return src * scale


SCALERS = {"identity": scale_identity, "amplification": scale_amplification, "attenuation": scale_attenuation,
           "linear": scale_linear, "inverse_linear": scale_inverse_linear}
BUILT = ("identity", "amplification", "attenuation")   # the set the fused kernel applies (reference default)
