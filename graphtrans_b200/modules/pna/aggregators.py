"""Names of the PNA aggregators (reference modules/pna/aggregators.py:37-44).  The arithmetic of
mean / max / min / std lives in the single-pass kernel gt_pna_reduce_* (csrc/pna.cu); the
reference's per-aggregator scatter functions have no standalone equivalent here."""
AGGREGATORS = {name: name for name in ("sum", "mean", "min", "max", "var", "std")}
BUILT = ("mean", "max", "min", "std")
