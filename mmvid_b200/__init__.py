"""mmvid_b200: B200-native (sm_100a) implementation of MMVID's token-level video generation hot path.

Public classes mirror the reference's Python API (same names / signatures / state-dict keys):
    BERT (mmvid_pytorch/dalle_bert.py:259), DALLE (mmvid_pytorch/dalle_artv.py:103),
    VQGanVAE1024 (mmvid_pytorch/vae.py:15), OpenAICLIPTransformer (mmvid_pytorch/transformers/clip_model.py:520)
All compute runs in libmmvid_b200.so (hand-written CUDA behind the C ABI in include/mmvid_b200.h); importing
the classes does not need a GPU, calling them does - there is no CPU fallback.
"""
__version__ = "0.1.0"

_LAZY = {
    "BERT": ("dalle_bert", "BERT"),
    "DALLE": ("dalle_artv", "DALLE"),
    "VQGanVAE1024": ("vae", "VQGanVAE1024"),
    "OpenAICLIPTransformer": ("transformer", "OpenAICLIPTransformer"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
