"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE. Runs only in the build container (the GPU box has no /root/reference):
    python oracle/gen_golden.py
The reference model classes are imported as they are through oracle/stubs (pure-torch
stand-ins for torch_geometric / torch_scatter / ogb, none of which is installed); each case is
run in fp64, train mode, dropout 0 (dropout RNG cannot be matched), and the
inputs, the reference-initialised state_dict, logits, loss, every parameter gradient and the
updated BatchNorm buffers are stored.  The fixtures pin oracle/graphtrans_oracle.py
(tests/test_oracle_golden.py) and, on the GPU, the CUDA path (tests/test_parity_gpu.py).
"""
import argparse
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
os.environ.setdefault("WANDB_MODE", "disabled")
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", ROOT]

from graphtrans_b200 import synth  # noqa: E402


def reference_model(args):
    """Build the reference model exactly as reference main.py:167-171 does."""
    import torch.nn as nn
    from models import MODELS
    from dataset.utils import ASTNodeEncoder
    from ogb.graphproppred.mol_encoder import AtomEncoder, BondEncoder

    ds = args.dataset
    if ds == "code2":
        node_encoder = ASTNodeEncoder(args.gnn_emb_dim, num_nodetypes=args.num_nodetypes,
                                      num_nodeattributes=args.num_nodeattrs, max_depth=20)
        edge_encoder_cls = lambda emb_dim: nn.Linear(2, emb_dim)  # noqa: E731
    elif ds in ("mol", "syn"):
        node_encoder = AtomEncoder(args.gnn_emb_dim)
        edge_encoder_cls = lambda emb_dim: BondEncoder(emb_dim=emb_dim)  # noqa: E731
    else:
        node_encoder = nn.Linear(37, args.gnn_emb_dim)

        def edge_encoder_cls(_):
            def zero(_):
                return 0
            return zero
    model_cls = MODELS[args.model_type]
    if ds != "code2" and model_cls.get_emb_dim(args) != args.gnn_emb_dim:     # reference dataset/mol.py:83, tud.py:65
        emb = model_cls.get_emb_dim(args)
        node_encoder = AtomEncoder(emb) if ds in ("mol", "syn") else nn.Linear(37, emb)
    return model_cls(num_tasks=args.num_tasks, args=args, node_encoder=node_encoder,
                     edge_encoder_cls=edge_encoder_cls)


def reference_loss(args):
    if args.dataset == "code2":
        from dataset.code import CodeUtil
        return CodeUtil.loss_fn(None)
    if args.dataset in ("mol", "syn"):
        from dataset.mol import MolUtil
        return MolUtil.loss_fn("binary classification")
    from dataset.tud import TUUtil
    return TUUtil.loss_fn(None)


def run_reference(args, batch, state_dict=None, dtype=torch.float64, seed=0):
    torch.manual_seed(seed)
    model = reference_model(args)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    init_sd = copy.deepcopy(model.state_dict())
    model = model.to(dtype)
    model.train()
    b = batch.clone()
    if torch.is_tensor(b.x) and b.x.is_floating_point():
        b.x = b.x.to(dtype)
    if torch.is_tensor(b.edge_attr) and b.edge_attr.is_floating_point():
        b.edge_attr = b.edge_attr.to(dtype)
    model.zero_grad()
    pred = model(b)
    loss = reference_loss(args)(pred, b)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in model.named_parameters()}
    bufs = {k: v.detach().clone() for k, v in model.named_buffers()}
    model.eval()
    with torch.no_grad():
        pred_eval = model(batch.clone() if dtype == torch.float32 else b.clone())
    det = lambda p: [t.detach() for t in p] if isinstance(p, list) else p.detach()  # noqa: E731
    return dict(init_sd=init_sd, logits=det(pred), loss=loss.detach(), grads=grads, buffers=bufs,
                logits_eval=det(pred_eval))


def edge_case_batch():
    """nci1-kind batch of hand-made edge cases: single-node graph, graph with
    no edges, duplicate edges, self loops, plus two random graphs."""
    rng = np.random.default_rng(7)
    graphs = [
        (1, []),                                            # single node
        (3, []),                                            # no edges
        (4, [(0, 1), (1, 0), (1, 2), (2, 1), (1, 2), (2, 1), (3, 3), (2, 3), (3, 2)]),  # dup + self loop
        (5, [(0, 1), (0, 2), (0, 3), (0, 4)]),              # directed star (out-deg != in-deg)
    ]
    for n in (9, 6):
        e = [(int(rng.integers(n)), int(rng.integers(n))) for _ in range(2 * n)]
        graphs.append((n, e + [(b, a) for a, b in e]))
    xs, src, dst, bidx, off = [], [], [], [], 0
    for g, (n, edges) in enumerate(graphs):
        x = np.zeros((n, 37), np.float32)
        x[np.arange(n), rng.integers(0, 37, size=n)] = 1
        xs.append(x)
        src += [a + off for a, _ in edges]
        dst += [b + off for _, b in edges]
        bidx += [g] * n
        off += n
    return synth.GraphBatch(
        x=torch.from_numpy(np.concatenate(xs)), edge_index=torch.tensor([src, dst], dtype=torch.long),
        edge_attr=None, batch=torch.tensor(bidx), num_graphs=len(graphs),
        y=torch.tensor([0, 1, 1, 0, 1, 0]))


def small_args(name, **kw):
    base = dict(gnn_emb_dim=32, d_model=32, dim_feedforward=64, gnn_num_layer=3, num_encoder_layers=2,
                gnn_dropout=0.0, transformer_dropout=0.0, num_nodetypes=98, num_nodeattrs=120)
    base.update(kw)
    return synth.make_args(name, **base)


def cases():
    out = {}
    a = small_args("nci1")
    out["gcn_plain_nci1"] = (a, synth.gen_nci1(6, seed=3))
    out["gcn_edgecases_nci1"] = (small_args("nci1"), edge_case_batch())
    a = small_args("code2", gnn_emb_dim=36, num_tasks=50, max_seq_len=3)
    out["gcn_virtual_cat_code2"] = (a, synth.gen_code2(5, seed=1, nmin=8, nmax=40, mu=3.0, sigma=0.5,
                                                       max_seq_len=3, num_nodeattrs=120, num_classes=50))
    a = small_args("molpcba", gnn_emb_dim=36, num_tasks=12)
    out["gin_virtual_cat_mol"] = (a, synth.gen_mol(7, seed=2, num_tasks=12))
    a = small_args("syn", gnn_emb_dim=32, num_tasks=12, gnn_JK="sum", gnn_num_layer=3)
    out["gin_plain_sum_syn"] = (a, synth.gen_syn(4, seed=4, num_tasks=12, nmin=5, nmax=20))
    a = small_args("code2", gnn_emb_dim=32, num_tasks=50, max_seq_len=3, gnn_JK="last", gnn_residual=True,
                   max_input_len=7)
    out["gcn_virtual_trunc_code2"] = (a, synth.gen_code2(5, seed=5, nmin=3, nmax=30, mu=2.5, sigma=0.6,
                                                         max_seq_len=3, num_nodeattrs=120, num_classes=50))
    b = synth.gen_code2(5, seed=6, nmin=8, nmax=40, mu=3.0, sigma=0.5, max_seq_len=3, num_nodeattrs=120,
                        num_classes=50)
    a = small_args("code2-pna", gnn_emb_dim=40, gnn_num_layer=2, num_tasks=50, max_seq_len=3,
                   deg=synth.in_degree_histogram(b, 800))
    out["pna_code2"] = (a, b)
    # baseline model families on the same kernels
    a = small_args("molpcba", model_type="gnn", gnn_emb_dim=36, num_tasks=12, gnn_JK="last", graph_pooling="mean")
    out["zgnn_gin_virtual_mean_mol"] = (a, synth.gen_mol(7, seed=8, num_tasks=12))
    a = small_args("nci1", model_type="gnn", gnn_JK="sum", graph_pooling="max")
    out["zgnn_gcn_max_nci1"] = (a, synth.gen_nci1(6, seed=9))
    b = synth.gen_mol(6, seed=10, num_tasks=12)
    a = small_args("molpcba", model_type="pna", gnn_emb_dim=40, gnn_num_layer=2, num_tasks=12, gnn_residual=True,
                   graph_pooling="sum", deg=synth.in_degree_histogram(b, 10))
    out["zpna_sum_mol"] = (a, b)
    a = small_args("molpcba", model_type="transformer", num_tasks=12, graph_pooling="cls")
    out["ztransformer_cls_mol"] = (a, synth.gen_mol(6, seed=11, num_tasks=12))
    a = small_args("nci1", model_type="transformer", graph_pooling="mean", max_input_len=9)
    out["ztransformer_mean_trunc_nci1"] = (a, synth.gen_nci1(5, seed=12))
    return out


def pad_batch_cases(out_dir):
    """integer/index fixture: the reference's own pad_batch / unpad_batch Python loops
    (reference modules/utils.py:5-53) on ragged batches incl. truncation and 1-node graphs."""
    from modules.utils import pad_batch, unpad_batch
    cases = []
    g = torch.Generator().manual_seed(0)
    for sizes in ([1], [3, 1, 5], [2, 9, 4, 1, 7], [12, 30, 1, 2, 2, 17], [5] * 4):
        for L in (1, 3, 7, 1000):
            for d in (4, 6):
                n = torch.tensor(sizes)
                batch = torch.repeat_interleave(torch.arange(len(sizes)), n)
                h = torch.randn(int(n.sum()), d, generator=g)
                padded, mask, num_nodes, masks, max_n = pad_batch(h, batch, L, get_mask=True)
                prev = torch.randn(int(n.sum()), d, generator=g)
                un = unpad_batch(padded, prev, num_nodes, masks, max_n)
                cases.append(dict(sizes=sizes, L=L, h=h, batch=batch, padded=padded, mask=mask,
                                  num_nodes=[int(v) for v in num_nodes], max_num_nodes=int(max_n), prev=prev,
                                  unpadded=un))
    torch.save(cases, os.path.join(out_dir, "_pad_batch_ref.pt"))
    print(f"_pad_batch_ref: {len(cases)} cases")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only-pad", action="store_true")
    ns = ap.parse_args()
    os.makedirs(ns.out, exist_ok=True)
    pad_batch_cases(ns.out)
    if ns.only_pad:
        return
    for name, (args, batch) in cases().items():
        res = run_reference(args, batch)
        fixture = dict(name=name, args=vars(args), batch={k: v for k, v in batch.__dict__.items()}, **res)
        path = os.path.join(ns.out, name + ".pt")
        torch.save(fixture, path)
        lg = res["logits"]
        lg0 = lg[0] if isinstance(lg, list) else lg
        print(f"{name}: loss={float(res['loss']):.10f} logits0[0,:3]={lg0[0, :3].tolist()} "
              f"params={sum(v.numel() for v in res['grads'].values())} size={os.path.getsize(path)/1e3:.0f}kB")


if __name__ == "__main__":
    main()
