"""CUDA-graph replay of the forward (conf.cuda_graph, a B200 extension): identical results to the eager launches."""
import pytest
import torch

from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_equals_eager(precision):
    torch.manual_seed(4)
    eager = LightGlue({"precision": precision, "filter_threshold": 0.1}).eval().to(DEV)
    graphed = LightGlue({"precision": precision, "filter_threshold": 0.1, "cuda_graph": True}).eval().to(DEV)
    graphed.load_state_dict(eager.state_dict())
    shapes = [(1, 300, 260, 41), (1, 300, 260, 42), (2, 256, 256, 43), (1, 300, 260, 44)]
    for B, n0, n1, seed in shapes:  # same signature three times (one capture, two replays), another one in between
        data = make_pairs(B, n0, n1, seed=seed, device=DEV)
        want, got = eager(data), graphed(data)
        assert set(want) == set(got)
        for k in want:
            assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype, k
            assert torch.equal(got[k], want[k]), k
    assert len(graphed._graphs) == 2
    # results are copies: a later replay does not overwrite what was returned earlier
    d1, d2 = make_pairs(1, 300, 260, seed=45, device=DEV), make_pairs(1, 300, 260, seed=46, device=DEV)
    o1 = graphed(d1)
    keep = o1["log_assignment"].clone()
    graphed(d2)
    assert torch.equal(o1["log_assignment"], keep)
    # new weights -> new capture (the packed weights are part of the signature)
    with torch.no_grad():
        graphed.log_assignment[-1].matchability.bias.add_(0.5)
        eager.log_assignment[-1].matchability.bias.add_(0.5)
    assert torch.equal(graphed(d1)["log_assignment"], eager(d1)["log_assignment"])
    # adaptive / training / per-pair counts fall back to eager launches
    d3 = dict(d1, num_keypoints0=torch.tensor([250]), num_keypoints1=torch.tensor([260]))
    assert torch.equal(graphed(d3)["matches0"], eager(d3)["matches0"])


def test_graph_survives_precision_switches():
    """A captured graph bakes in device pointers into the packed weights.  Switching the precision replaces the
    module's pack; switching back re-creates an identical signature and must replay a graph whose weights are still
    alive (the graph entry owns them), not freed memory."""
    torch.manual_seed(8)
    model = LightGlue({"precision": "fp32", "filter_threshold": 0.1, "cuda_graph": True}).eval().to(DEV)
    eager = LightGlue({"precision": "fp32", "filter_threshold": 0.1}).eval().to(DEV)
    eager.load_state_dict(model.state_dict())
    data = make_pairs(1, 300, 260, seed=47, device=DEV)
    want = {}
    for prec in ("fp32", "bf16"):
        eager.conf.precision = prec
        want[prec] = eager(data)["log_assignment"].clone()
    for rnd, prec in enumerate(["fp32", "bf16", "fp32", "bf16", "fp32"]):
        model.conf.precision = prec
        # churn the allocator between switches so that a freed pack would be reused by something else
        junk = [torch.randn(1 << 20, device=DEV) for _ in range(8)]
        got = model(data)["log_assignment"]
        assert torch.equal(got, want[prec]), f"round {rnd} ({prec})"
        del junk
    assert len(model._graphs) == 2  # one capture per precision, reused after the switches
    # static outputs: no copies, valid until the next call
    model.conf.graph_static_outputs = True
    o = model(data)
    assert torch.equal(o["log_assignment"], want["fp32"])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("mode", ["depth+width", "width"])
def test_adaptive_graph_replay_equals_eager(precision, mode):
    """Adaptive depth / width under conf.cuda_graph: the transformer stack is replayed from a graph (all layers; the
    kernels skip finished pairs from the device-side counts), the data-dependent tail runs eagerly.  Same results as
    the eager adaptive forward, including the pruned log_assignment shape, prune layers and matches."""
    conf = {"precision": precision, "filter_threshold": 0.1, "width_confidence": 0.99}
    if mode == "depth+width":
        conf["depth_confidence"] = 0.95
    torch.manual_seed(6)
    eager = LightGlue(conf).eval()
    sd = eager.state_dict()
    for i in range(8):  # heads that actually prune and (depth) exit at layer 4, as in tools/adaptive_bench.py
        sd[f"token_confidence.{i}.token.0.bias"].fill_(3.0 if i >= 4 else -3.0)
        sd[f"log_assignment.{i}.matchability.bias"].fill_(-4.5 if i % 2 == 0 else 0.0)
    eager = eager.to(DEV)
    graphed = LightGlue({**conf, "cuda_graph": True}).eval().to(DEV)
    graphed.load_state_dict(eager.state_dict())
    for B, n0, n1, seed in [(1, 512, 480, 51), (1, 512, 480, 52), (3, 384, 384, 53), (1, 512, 480, 54)]:
        data = make_pairs(B, n0, n1, seed=seed, device=DEV)
        want, got = eager(data), graphed(data)
        assert set(want) == set(got)
        for k in want:
            assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype, k
            assert torch.equal(got[k], want[k]), k
        assert float(want["prune0"].float().mean()) < 9.0  # something was pruned / exited early
    assert len(graphed._graphs) == 2
