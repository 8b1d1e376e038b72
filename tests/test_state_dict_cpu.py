"""Drop-in contract: the product models expose exactly the reference's
state_dict keys / shapes / dtypes, so checkpoints interoperate both ways.  The golden fixtures
hold state_dicts initialised by the UNMODIFIED reference.  CPU only (no kernels are called)."""
import pytest
import torch

from graphtrans_b200 import factory, synth
from tests.helpers import GOLDEN_CASES, load_golden


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_state_dict_matches_reference(name):
    fx = load_golden(name)
    model = factory.build_model(fx["args"])
    ref = fx["init_sd"]
    own = model.state_dict()
    assert list(own.keys()) == list(ref.keys()) or set(own.keys()) == set(ref.keys())
    for k, v in ref.items():
        assert own[k].shape == v.shape, k
        assert own[k].dtype == v.dtype, k
    model.load_state_dict(ref, strict=True)


def test_run_names_follow_the_reference():
    from graphtrans_b200.models import MODELS
    a = synth.make_args("code2")
    assert MODELS["gnn-transformer"].name(a) == (
        "gnn-transformer-pooling=cls-norm_input+gcn-virtual-JK=cat-enc_layer=4-enc_layer_masked=0-d=256-act=relu"
        "-tdrop=0.3-gdrop=0.0-postnorm")
    a = synth.make_args("code2-pna")
    assert MODELS["pna-transformer"].name(a).startswith("pna-transformer-pooling=cls-norm_input+gcn-JK=last-enc_layer=4")


def test_flags_and_defaults():
    import argparse
    from graphtrans_b200.models import get_model_and_parser
    p = argparse.ArgumentParser()
    p.add_argument("--graph_pooling", default="cls")
    ns = argparse.Namespace(model_type="gnn-transformer")
    get_model_and_parser(ns, p)
    a = p.parse_args([])
    assert (a.d_model, a.nhead, a.dim_feedforward, a.transformer_dropout, a.num_encoder_layers) == (128, 4, 512, 0.3, 4)
    assert a.max_input_len == 1000 and a.num_encoder_layers_masked == 0 and a.transformer_norm_input is False
    p = argparse.ArgumentParser()
    p.add_argument("--gnn_residual", default=False)
    p.add_argument("--gnn_dropout", default=0.0)
    p.add_argument("--gnn_emb_dim", default=300)
    p.add_argument("--gnn_num_layer", default=5)
    get_model_and_parser(argparse.Namespace(model_type="pna-transformer"), p)
    a = p.parse_args([])
    assert a.gnn_residual is True and a.gnn_dropout == 0.3 and a.gnn_emb_dim == 70 and a.gnn_num_layer == 4
    assert a.aggregators == ["mean", "max", "min", "std"]
