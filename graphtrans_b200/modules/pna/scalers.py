"""Names of the PNA degree scalers (reference modules/pna/scalers.py:34-40); identity /
amplification / attenuation are folded into gt_pna_reduce_* (csrc/pna.cu)."""
SCALERS = {name: name for name in ("identity", "amplification", "attenuation", "linear", "inverse_linear")}
BUILT = ("identity", "amplification", "attenuation")
