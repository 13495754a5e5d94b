"""Batched pair driver: host planning / collation on CPU, end-to-end equality on the GPU."""
import pytest
import torch

from glue_factory_colon_b200.driver import BatchedPairMatcher, collate_pairs, crop_log_assignment, plan_batches
from glue_factory_colon_b200.synthetic import make_pairs


def _pair(n0, n1, seed, dim=256):
    d = make_pairs(1, n0, n1, seed=seed, dim=dim)
    return {
        "keypoints0": d["keypoints0"][0], "keypoints1": d["keypoints1"][0],
        "descriptors0": d["descriptors0"][0], "descriptors1": d["descriptors1"][0],
        "image_size0": d["view0"]["image_size"][0], "image_size1": d["view1"]["image_size"][0],
    }


def test_plan_batches_covers_every_pair_once_in_order_and_respects_budgets():
    g = torch.Generator().manual_seed(0)
    counts = [(int(a), int(b)) for a, b in torch.randint(1, 3000, (200, 2), generator=g)]
    plan = plan_batches(counts, max_pairs=16, max_tokens=16 * 2 * 2048, bucket=128)
    flat = [i for b in plan for i in b]
    assert flat == list(range(200))
    for b in plan:
        assert 1 <= len(b) <= 16
        n0 = max(-(-counts[i][0] // 128) * 128 for i in b)
        n1 = max(-(-counts[i][1] // 128) * 128 for i in b)
        assert len(b) == 1 or len(b) * (n0 + n1) <= 16 * 2 * 2048
    # an oversized pair still gets (its own) batch
    assert plan_batches([(9000, 9000), (10, 10)], max_pairs=4, max_tokens=4096) == [[0], [1]]
    assert plan_batches([], 4) == []
    with pytest.raises(ValueError):
        plan_batches([(1, 1)], max_pairs=0)


def test_plan_batches_by_size_groups_similar_pairs():
    g = torch.Generator().manual_seed(1)
    counts = [(int(a), int(b)) for a, b in torch.randint(200, 4097, (256, 2), generator=g)]

    def padded_cost(plan):  # attention work of the padded batches: B * (N0p^2 + N1p^2 + 1.5 N0p N1p)
        tot = 0
        for b in plan:
            n0 = max(-(-counts[i][0] // 128) * 128 for i in b)
            n1 = max(-(-counts[i][1] // 128) * 128 for i in b)
            tot += len(b) * (n0 * n0 + n1 * n1 + 1.5 * n0 * n1)
        return tot

    kw = dict(max_pairs=32, max_tokens=32 * 2 * 4096, bucket=128)
    plain, sized = plan_batches(counts, **kw), plan_batches(counts, by_size=True, **kw)
    assert sorted(i for b in sized for i in b) == list(range(256))  # every pair exactly once
    for b in sized:
        n0 = max(-(-counts[i][0] // 128) * 128 for i in b)
        n1 = max(-(-counts[i][1] // 128) * 128 for i in b)
        assert 1 <= len(b) <= 32 and (len(b) == 1 or len(b) * (n0 + n1) <= 32 * 2 * 4096)
    assert padded_cost(sized) < 0.8 * padded_cost(plain)


def test_collate_pads_with_zeros_and_records_counts():
    pairs = [_pair(100, 257, 1), _pair(300, 5, 2)]
    out = collate_pairs(pairs, bucket=128)
    assert out["keypoints0"].shape == (2, 384, 2) and out["keypoints1"].shape == (2, 384, 2)
    assert out["descriptors0"].shape == (2, 384, 256)
    assert out["num_keypoints0"].tolist() == [100, 300] and out["num_keypoints1"].tolist() == [257, 5]
    assert torch.equal(out["descriptors1"][0, :257], pairs[0]["descriptors1"])
    assert (out["descriptors1"][1, 5:] == 0).all() and (out["keypoints0"][0, 100:] == 0).all()
    assert out["view0"]["image_size"].shape == (2, 2)
    with pytest.raises(ValueError):
        bad = dict(pairs[1]); bad.pop("image_size0")
        collate_pairs([pairs[0], bad])


def test_crop_log_assignment_moves_the_dustbins():
    s = torch.arange(6 * 7, dtype=torch.float32).reshape(6, 7)
    c = crop_log_assignment(s, 2, 3)
    assert c.shape == (3, 4)
    assert torch.equal(c[:2, :3], s[:2, :3]) and torch.equal(c[:2, 3], s[:2, 6]) and torch.equal(c[2, :3], s[5, :3])
    assert c[2, 3] == s[5, 6]


@pytest.mark.gpu
def test_driver_equals_one_pair_at_a_time_fp32():
    """The reference call site runs one pair per call; the driver's padded batches must give the same matches."""
    from glue_factory_colon_b200 import LightGlue

    torch.manual_seed(0)
    model = LightGlue({"filter_threshold": 0.1, "precision": "fp32"}).eval().cuda()
    sizes = [(130, 250), (512, 300), (64, 700), (257, 129), (400, 400), (33, 90), (640, 128)]
    pairs = [_pair(a, b, 10 + i) for i, (a, b) in enumerate(sizes)]
    drv = BatchedPairMatcher(model, max_pairs=3, max_tokens=3 * 2 * 768, return_log_assignment=True)
    got = list(drv.match(iter(pairs)))
    assert len(got) == len(pairs)
    for p, r, (a, b) in zip(pairs, got, sizes):
        single = {
            "keypoints0": p["keypoints0"][None].cuda(), "keypoints1": p["keypoints1"][None].cuda(),
            "descriptors0": p["descriptors0"][None].cuda(), "descriptors1": p["descriptors1"][None].cuda(),
            "view0": {"image_size": p["image_size0"][None].cuda()}, "view1": {"image_size": p["image_size1"][None].cuda()},
        }
        ref = model(single)
        assert r["matches0"].shape == (a,) and r["matches1"].shape == (b,)
        torch.testing.assert_close(r["log_assignment"].cpu(), ref["log_assignment"][0].cpu(), atol=2e-4, rtol=1e-4)
        agree = (r["matches0"] == ref["matches0"][0].cpu()).float().mean()
        assert agree >= 0.995, agree  # identical up to fp32 summation-order ties at the threshold
        torch.testing.assert_close(r["matching_scores1"], ref["matching_scores1"][0].cpu(), atol=1e-3, rtol=1e-3)


@pytest.mark.gpu
def test_driver_bf16_streams_many_batches_in_order():
    from glue_factory_colon_b200 import LightGlue

    torch.manual_seed(0)
    model = LightGlue({"filter_threshold": 0.1, "precision": "bf16"}).eval().cuda()
    g = torch.Generator().manual_seed(3)
    sizes = [(int(a), int(b)) for a, b in torch.randint(64, 900, (40, 2), generator=g)]
    pairs = [_pair(a, b, 100 + i) for i, (a, b) in enumerate(sizes)]
    drv = BatchedPairMatcher(model, max_pairs=8, window=16)
    got = list(drv.match(pairs))
    assert [tuple(map(len, (r["matches0"], r["matches1"]))) for r in got] == sizes
    n_valid = 0
    for r, (a, b) in zip(got, sizes):
        m0 = r["matches0"]
        n_valid += int((m0 >= 0).sum())
        assert ((m0 >= -1) & (m0 < b)).all()
        valid = m0 >= 0
        assert (r["matches1"][m0[valid]] == torch.nonzero(valid).squeeze(-1)).all()  # mutual consistency
    assert n_valid > 0  # (random-init weights: only some pairs pass the 0.1 threshold)
