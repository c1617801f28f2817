"""Shared case table for the golden fixtures (used by gen_golden.py here and by the tests everywhere).

A case fixes: model hyper-parameters, the seed of the synthetic weights (`mmvid_b200.synth`) and the seeded
inputs.  Fixtures hold only reference OUTPUTS; weights/inputs are regenerated from the seeds.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))

# name -> config.  `dim` must be a multiple of 64 (CLIP: heads = width // 64).
BERT_CASES = {
    # tiny: fast everywhere; exercises cvae + visual control + preserve
    "bert_tiny": dict(dim=128, layers=2, text_seq_len=6, vocab=64, num_visuals=1, num_targets=2, image_size=32,
                      cvae=True, seed=11, batch=2),
    "bert_tiny_nov": dict(dim=128, layers=2, text_seq_len=5, vocab=64, num_visuals=0, num_targets=3, image_size=32,
                          cvae=False, seed=12, batch=2),
    # full width, Shape B (reference scripts: 128 px, text 50) and Shape A (BASELINE: seq ~2k)
    "bert_shapeB": dict(dim=768, layers=12, text_seq_len=50, vocab=49408, num_visuals=0, num_targets=8,
                        image_size=128, cvae=False, seed=21, batch=1),
    "bert_shapeA": dict(dim=768, layers=12, text_seq_len=64, vocab=49408, num_visuals=0, num_targets=8,
                        image_size=256, cvae=False, seed=22, batch=1),
    # BASELINE config 4 (text + mask conditioning, scripts/mmvoxceleb/text_and_mask): one cVAE-encoded visual-control frame
    # whose token grid is windowed by vc_mode='mask_8x8' (face_mode given => deterministic strategy 3), S = 629
    "bert_shapeB_vis": dict(dim=768, layers=12, text_seq_len=50, vocab=49408, num_visuals=1, num_targets=8,
                            image_size=128, cvae=True, seed=23, batch=2, vc_mode="mask_8x8", face_mode="mouth"),
}

ARTV_CASES = {
    "artv_tiny": dict(dim=128, layers=2, text_seq_len=6, vocab=64, num_visuals=1, num_targets=2, image_size=32,
                      seed=31, batch=2),
}

VAE_CASES = {
    # BASELINE config 1: 1 frame 64x64 encode -> quantize -> decode
    "vae_64": dict(image_size=64, batch=1, seed=41),
    "vae_32": dict(image_size=32, batch=3, seed=42),
    "vae_128": dict(image_size=128, batch=2, seed=43),
}

TRANSFORMER_CASES = {
    "tfm_small": dict(dim=128, layers=2, seq=37, batch=2, mask="mask_prev", index=(9, 10), seed=51),
    "tfm_causal": dict(dim=128, layers=2, seq=40, batch=1, mask="causal", index=(), seed=52),
    "tfm_wide": dict(dim=768, layers=2, seq=150, batch=1, mask="mask_prev", index=(20, 21), seed=53),
}


def fixture_path(name):
    return os.path.join(GOLDEN_DIR, name + ".pt")


def codebook_std():
    """Std of the synthetic VQ codebook.  Pre-quant latents from the synthetic encoder have O(0.3) entries;
    a N(0, 0.3^2) codebook gives well separated nearest neighbours."""
    return 0.3
