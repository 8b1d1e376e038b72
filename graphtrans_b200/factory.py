"""Model / loss construction exactly as the reference's driver does it (reference main.py:126-171
with dataset/code.py:117-133, dataset/mol.py:83-85, dataset/tud.py:65-73), for the synthetic
dataset kinds of graphtrans_b200.synth.  The losses are the caller-side functions of the
reference's dataset adapters (dataset/code.py:39-45, dataset/mol.py:24-31, dataset/tud.py:25-27)
operating on the fp32 logits the model returns."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import synth
from .encoders import ASTNodeEncoder, AtomEncoder, BondEncoder
from .models import MODELS


def build_model(args):
    ds = args.dataset
    model_cls = MODELS[args.model_type]
    emb = model_cls.get_emb_dim(args)        # reference dataset/mol.py:83: the node encoder takes the MODEL's width
    if ds == "code2":
        node_encoder = ASTNodeEncoder(emb, num_nodetypes=getattr(args, "num_nodetypes", synth.CODE2_NUM_NODETYPES),
                                      num_nodeattributes=getattr(args, "num_nodeattrs", synth.CODE2_NUM_NODEATTRS),
                                      max_depth=synth.CODE2_MAX_DEPTH)
        edge_encoder_cls = lambda emb_dim: nn.Linear(2, emb_dim)  # noqa: E731
    elif ds in ("mol", "syn"):
        node_encoder = AtomEncoder(emb)
        edge_encoder_cls = lambda emb_dim: BondEncoder(emb_dim=emb_dim)  # noqa: E731
    elif ds == "nci1":
        node_encoder = nn.Linear(37, emb)

        def edge_encoder_cls(_):
            def zero(_):
                return 0
            return zero
    else:
        raise ValueError(ds)
    return model_cls(num_tasks=args.num_tasks, args=args, node_encoder=node_encoder, edge_encoder_cls=edge_encoder_cls)


def predict_fn(args):
    """eval read-out of the reference's dataset adapters on the logits the model returns: Code2 -> int64 [B, max_seq_len]
    token ids (reference dataset/code.py:62-66: argmax per head, ONE launch over the stacked logits here); TU -> argmax
    class (dataset/tud.py:37); mol -> the logits themselves (dataset/mol.py:47-49 hands them to the OGB evaluator)"""
    from . import ops
    ds = args.dataset
    if ds == "code2":
        def predict(pred_list):
            st = getattr(pred_list, "stacked", None)
            if st is not None:
                y, rp, n_cls = st
                return ops.argmax_rows(y.view(-1, rp), n_cols=n_cls).view(y.shape[0], -1)
            return torch.stack([ops.argmax_rows(p) for p in pred_list], dim=1)
        return predict
    if ds == "nci1":
        return lambda pred: ops.argmax_rows(pred)
    return lambda pred: pred


def loss_fn(args):
    """the reference's calc_loss(pred, batch, m=1.0) per dataset kind; on CUDA tensors the fused device-resident
    kernels (gt_ce_*, gt_bce_masked_*) are used, otherwise the reference's torch formulation"""
    from . import ops
    ds = args.dataset
    if ds == "code2":
        def calc_loss(pred_list, batch, m=1.0):
            st = getattr(pred_list, "stacked", None)
            if st is not None and batch.y_arr.is_contiguous() and batch.y_arr.shape[1] == len(pred_list):
                # heads stacked as [B, H, Np]: mean CE over the B * H rows == (1/H) sum_h mean_b CE_h, one launch
                y, rp, n_cls = st
                return ops.cross_entropy_mean(y.view(-1, rp), batch.y_arr.view(-1), n_cols=n_cls) / m
            loss = 0
            for i in range(len(pred_list)):
                if pred_list[i].is_cuda:
                    loss = loss + ops.cross_entropy_mean(pred_list[i], batch.y_arr[:, i])
                else:
                    loss = loss + F.cross_entropy(pred_list[i].to(torch.float32), batch.y_arr[:, i])
            return loss / len(pred_list) / m
        return calc_loss
    if ds in ("mol", "syn"):
        def calc_loss(pred, batch, m=1.0):
            # mean over labelled (non-NaN) entries as dataset/mol.py:24-31 (the reference's boolean gather
            # pred[is_labeled] forces a device->host sync; the kernel counts and sums on the device)
            if pred.is_cuda:
                return ops.bce_with_logits_masked_mean(pred, batch.y) / m
            is_labeled = batch.y == batch.y
            return F.binary_cross_entropy_with_logits(pred.to(torch.float32)[is_labeled],
                                                      batch.y.to(torch.float32)[is_labeled]) / m
        return calc_loss
    if ds == "nci1":
        def calc_loss(pred, batch, m=1.0):
            if pred.is_cuda:
                return ops.cross_entropy_mean(pred, batch.y) / m
            return F.cross_entropy(pred, batch.y) / m
        return calc_loss
    raise ValueError(ds)
