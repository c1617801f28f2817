"""GPU: CLIP scoring on the library's kernels (mmvid_b200/clip_score.py) against the reference CLIP's outputs
(tests/golden/clip_small.pt, produced by the unmodified clip_model.CLIP on CPU fp32, tests/golden/gen_clip_golden.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


class _Tok:
    """tokenizer stand-in that returns the fixture's ids (the BPE table is not on the GPU box)"""

    def __init__(self, text):
        self.text = text

    def tokenize(self, description, context_length, truncate_text=False):
        assert context_length == self.text.shape[1] and truncate_text
        return self.text


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32", 1e-3), ("fp16", 1e-3)])
def test_clip_features_and_similarity_match_reference(prec, tol):
    from mmvid_b200.clip_score import CLIP, clip_similarity
    fx = torch.load(os.path.join(HERE, "golden", "clip_small.pt"))
    model = CLIP(**fx["cfg"], precision=prec)
    missing = model.load_state_dict(fx["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert set(model.state_dict().keys()) == set(fx["state_dict"].keys()), "same state-dict keys as clip_model.CLIP"
    model = model.cuda().eval()
    image, text = fx["image"].cuda(), fx["text"].cuda()
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device="cuda")
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device="cuda")
    fi = model.encode_image((image - mean[:, None, None]) / std[:, None, None]).cpu()
    ft = model.encode_text(text).cpu()
    ei = float((fi - fx["image_features"]).norm() / fx["image_features"].norm())
    et = float((ft - fx["text_features"]).norm() / fx["text_features"].norm())
    sim = clip_similarity(model, _Tok(text), image, ["a", "b", "c"])
    es = float(abs(torch.from_numpy(sim) - fx["similarity"]).max())
    li, _ = model(((image - mean[:, None, None]) / std[:, None, None]), text)
    el = float((li.cpu() - fx["logits_per_image"]).norm() / fx["logits_per_image"].norm())
    print(f"clip_small {prec}: image features relerr {ei:.2e}, text features {et:.2e}, similarity max abs err {es:.2e}, logits {el:.2e}")
    assert ei < tol and et < tol and es < tol and el < 5 * tol
    # a resolution other than the model's goes through the same nearest resize as utils/utils.py:66-67
    big = torch.nn.functional.interpolate(image, (192, 192))
    assert abs(torch.from_numpy(clip_similarity(model, _Tok(text), big, ["a", "b", "c"])) - torch.from_numpy(sim)).max() < 1e-6
