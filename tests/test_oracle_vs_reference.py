"""CPU, build container only: re-check the oracle against the LIVE reference (skipped where /root/reference is absent)."""
import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not present")


def test_transformer_oracle_vs_live_reference():
    from cases import TRANSFORMER_CASES
    from helpers import relerr
    from mmvid_b200 import synth
    from oracle import mmvid_oracle as O
    ref_shims.install()
    from mmvid_pytorch.transformers.clip_model import OpenAICLIPTransformer
    cfg = TRANSFORMER_CASES["tfm_small"]
    with ref_shims.fake_clip_checkpoint(synth.clip_checkpoint_state_dict(cfg["dim"], cfg["layers"], seed=1)):
        m = OpenAICLIPTransformer(cfg["seq"], "openai_clip_visual", model_path="none", causal=True,
                                  mask_type="mask_prev", mask_kwargs={"index": list(cfg["index"])})
    sd = synth.fill_state_dict(m, 99)
    m.load_state_dict(sd)
    x = torch.randn(2, cfg["seq"], cfg["dim"])
    with torch.no_grad():
        y_ref = m.eval()(x)
        y = O.transformer_forward(x, sd, "transformer.", O.build_attention_mask(cfg["seq"], "mask_prev", cfg["index"]))
    assert relerr(y, y_ref) < 2e-6
