"""BASELINE config 1 fixture (tests/golden/c1_boat.pt): the UNMODIFIED reference LightGlue
(/root/reference/gluefactory/models/matchers/lightglue.py) on the reference's own demo pair
assets/boat1.png / boat2.png (850 x 680), 1024 keypoints per image, conf values of
configs/superpoint+lightglue-official.yaml:10-13 (filter_threshold 0.1, adaptive depth / width off).

Substitutions, as SURVEY.md 8(c) "Consequence for config 1" prescribes (no network: neither superpoint_v1.pth nor the
official LightGlue weights exist here):
  * keypoints  = OpenCV SIFT detections on the two images (strongest 1024 each) -- real locations;
  * descriptors = the SIFT descriptors of those points (integers 0..255, stored as uint8) lifted to 256-d by a seeded
    Gaussian projection and L2-normalised -- real appearance, so the pair has true correspondences;
  * weights    = the reference constructor under torch.manual_seed(21), with the last MatchAssignment made sharp
    (final_proj = 2 I, matchability bias +4) so that filter_threshold 0.1 keeps a few hundred matches instead of none.

Run in the build container only (needs /root/reference and cv2):  python oracle/make_golden_c1.py
The fixture stores the inputs, the reference's matches / scores, and log_assignment sub-sampled 4 x 4 plus its
dustbin row and column (the full 1025 x 1025 matrix is compared against the live CPU oracle in the GPU test).
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle" / "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))

SEED, PROJ_SEED, N = 21, 22, 1024
CONF = {"filter_threshold": 0.1, "depth_confidence": -1, "width_confidence": -1}


def overrides():
    return {
        "log_assignment.8.final_proj.weight": 2.0 * torch.eye(256),
        "log_assignment.8.final_proj.bias": torch.zeros(256),
        "log_assignment.8.matchability.bias": torch.tensor([4.0]),
    }


def lift_descriptors(sift_u8: torch.Tensor) -> torch.Tensor:
    """[n,128] uint8 SIFT -> [n,256] unit fp32 (seeded projection; also used by the test to rebuild the inputs)."""
    g = torch.Generator().manual_seed(PROJ_SEED)
    P = torch.randn(128, 256, generator=g) / 128 ** 0.5
    d = torch.nn.functional.normalize(sift_u8.float(), dim=-1)
    return torch.nn.functional.normalize(d @ P, dim=-1)


def detect(path):
    import cv2
    import numpy as np

    im = cv2.imread(str(path), cv2.IMREAD_GRAYSCALE)
    kps, desc = cv2.SIFT_create(nfeatures=N).detectAndCompute(im, None)
    order = np.argsort([-k.response for k in kps], kind="stable")[:N]
    xy = torch.tensor([[kps[i].pt[0], kps[i].pt[1]] for i in order], dtype=torch.float32)
    d = torch.from_numpy(desc[order]).round().clamp(0, 255).to(torch.uint8)
    return xy, d, (float(im.shape[1]), float(im.shape[0]))


def main():
    from gluefactory.models import get_model

    k0, s0, wh0 = detect("/root/reference/assets/boat1.png")
    k1, s1, wh1 = detect("/root/reference/assets/boat2.png")
    assert k0.shape == (N, 2) and k1.shape == (N, 2)
    torch.manual_seed(SEED)
    model = get_model("matchers.lightglue")(dict(CONF)).eval()
    sd = model.state_dict()
    for k, v in overrides().items():
        sd[k].copy_(v)
    data = {
        "keypoints0": k0[None], "keypoints1": k1[None],
        "descriptors0": lift_descriptors(s0)[None], "descriptors1": lift_descriptors(s1)[None],
        "view0": {"image_size": torch.tensor([wh0])}, "view1": {"image_size": torch.tensor([wh1])},
    }
    with torch.no_grad():
        out = model(data)
    la = out["log_assignment"][0]
    fx = {
        "conf": CONF, "seed": SEED, "keypoints0": k0, "keypoints1": k1, "sift0": s0, "sift1": s1,
        "image_size0": wh0, "image_size1": wh1,
        "fingerprint": float(sum(v.double().abs().sum() for v in model.state_dict().values())),
        "matches0": out["matches0"][0].clone(), "matches1": out["matches1"][0].clone(),
        "matching_scores0": out["matching_scores0"][0].clone(), "matching_scores1": out["matching_scores1"][0].clone(),
        "la_sub": la[::4, ::4].clone(), "la_dust_col": la[:, -1].clone(), "la_dust_row": la[-1, :].clone(),
    }
    torch.save(fx, ROOT / "tests" / "golden" / "c1_boat.pt")
    print("c1_boat: valid matches", int((out["matches0"] > -1).sum()), "of", N,
          "| la range", float(la.min()), float(la.max()))


if __name__ == "__main__":
    main()
