"""Weight wire format helpers: containers and prefixes only, CPU."""
import collections

import pytest
import torch

from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.weights import (export_plain_state_dict, load_matcher_weights, strip_prefixes,
                                              to_pipeline_state_dict, unwrap_checkpoint)


def _model(seed):
    torch.manual_seed(seed)
    return LightGlue({"n_layers": 2})


def test_plain_export_round_trip(tmp_path):
    a, b = _model(0), _model(1)
    path = tmp_path / "lg.pt"
    plain = export_plain_state_dict(a, path)
    assert type(plain) is dict and all(v.device.type == "cpu" for v in plain.values())
    loaded = torch.load(path, map_location="cpu", weights_only=True)
    assert type(loaded) is dict and list(loaded) == list(a.state_dict())
    res = load_matcher_weights(b, path)
    assert not res.missing_keys and not res.unexpected_keys
    for (k, va), vb in zip(a.state_dict().items(), b.state_dict().values()):
        assert torch.equal(va, vb), k


def test_checkpoint_containers_and_prefixes():
    a, b = _model(2), _model(3)
    sd = a.state_dict()
    pipeline = collections.OrderedDict((f"module.matcher.{k}", v) for k, v in sd.items())
    pipeline["module.extractor.conv1.weight"] = torch.zeros(3)  # another pipeline member: ignored
    ckpt = {"model": pipeline, "epoch": 7}
    assert set(strip_prefixes(unwrap_checkpoint(ckpt))) == set(sd)
    load_matcher_weights(b, ckpt)
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    back = to_pipeline_state_dict(pipeline)
    assert set(back) == {f"matcher.{k}" for k in sd}
    with pytest.raises(ValueError):
        unwrap_checkpoint({"model": 3})
    with pytest.raises(ValueError):
        unwrap_checkpoint([1, 2])


def test_training_path_covers_every_parameter():
    """train.transformer_params (what TransformerFn differentiates) + the MatchAssignment and token-confidence heads
    (AssignFn / torch in LightGlue.loss) = every parameter of the module, each exactly once."""
    import torch

    from glue_factory_colon_b200 import LightGlue
    from glue_factory_colon_b200.train import transformer_params

    for conf in ({}, {"input_dim": 128, "add_scale_ori": True, "n_layers": 3}):
        model = LightGlue(conf)
        covered = [p for _, p in transformer_params(model)]
        covered += [p for a in model.log_assignment for p in a.parameters()]
        covered += [p for t in model.token_confidence for p in t.parameters()]
        ids = [id(p) for p in covered]
        assert len(ids) == len(set(ids))
        assert set(ids) == {id(p) for p in model.parameters()}
        keys = [k for k, _ in transformer_params(model)]
        assert len(keys) == len(set(keys))
    # the packed Wqkv order used by the backward pass is a permutation of the reference's rows
    from glue_factory_colon_b200.train import _PERM

    assert sorted(_PERM.tolist()) == list(range(768))
    assert int(_PERM[0]) == 0 and int(_PERM[1]) == 3 and int(_PERM[256]) == 1  # part*256 + head*64 + d <- head*192 + d*3 + part
