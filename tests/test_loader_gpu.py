"""DevicePrefetcher: batches arrive on the device unchanged, one copy ahead on a side stream, and feed the model."""
import pytest
import torch

from graphtrans_b200 import factory, loader, ops, synth

pytestmark = pytest.mark.gpu


def test_prefetched_batches_equal_their_host_sources_and_train():
    args = synth.make_args("molpcba", gnn_dropout=0.0, transformer_dropout=0.0)
    host = [synth.make_batch(args, B=6, seed=s) for s in range(4)]
    ops.set_precision("fp32")
    torch.manual_seed(0)
    model = factory.build_model(args).cuda().train()
    lossf = factory.loss_fn(args)
    seen = 0
    for hb, db in zip(host, loader.DevicePrefetcher(host, "cuda")):
        for k, v in hb.__dict__.items():
            w = getattr(db, k)
            if torch.is_tensor(v):
                assert w.is_cuda and torch.equal(torch.nan_to_num(w.cpu().float()), torch.nan_to_num(v.float())), k
            else:
                assert w == v, k
        loss = lossf(model(db), db)
        loss.backward()
        assert torch.isfinite(loss)
        seen += 1
    assert seen == len(host)
    assert list(loader.DevicePrefetcher([], "cuda")) == []
