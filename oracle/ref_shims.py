"""Import shims that let the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Used by `tests/golden/gen_golden.py` (fixture generator) and by the
`-m "not gpu"` tests that pin `oracle/mmvid_oracle.py` against the reference itself.  `/root/reference`
does not exist on the GPU box; everything here degrades to "reference unavailable" there.

What is shimmed - each is a *missing third-party module*, never reference code:
  1. axial_positional_embedding.AxialPositionalEmbedding (pip, unpinned in requirements.txt:2;
     call sites dalle_bert.py:8,326  dalle_artv.py:8,141  modules.py:4,24).  Restated from the published
     lucidrains package (v0.2.x): one nn.Parameter `weights_i` per axis, shape (1, 1.., n_i, ..1, dim),
     N(0,1) init; forward = sum over axes of the broadcast tables, flattened, sliced to the input length.
     "parity unpinned" at this boundary: no reference test pins it.
  2. pytorch_lightning.LightningModule -> nn.Module   (taming/models/vqgan.py:3,16)
  3. omegaconf.OmegaConf.load -> yaml.safe_load attr-dict; lossconfig -> torch.nn.Identity
     (vae.py:22-26; vqgan.py:35 would build LPIPS, which downloads VGG16)
  4. torchvision.io.write_video -> noop  (utils/utils.py:8)
  5. torch.jit.load(ViT-B-32.pt) -> seeded stand-in exposing .state_dict() of CLIP-shaped weights
     (clip_model.py:535-536); the checkpoint is not available offline.
"""
import contextlib
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("MMVID_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "mmvid_pytorch"))


class _AxialPositionalEmbedding(nn.Module):
    """Parameters are registered on the module itself as `weights_0`, `weights_1`, ... exactly like the
    published package (released MMVID checkpoints carry `target_pos_emb.weights_{0,1,2}`)."""

    def __init__(self, dim, axial_shape, axial_dims=None):
        super().__init__()
        self.dim = dim
        self.shape = tuple(axial_shape)
        self.max_seq_len = 1
        for s in self.shape:
            self.max_seq_len *= s
        for ind, n in enumerate(self.shape):
            ax_shape = [1] * len(self.shape)
            ax_shape[ind] = n
            self.register_parameter(f"weights_{ind}", nn.Parameter(torch.zeros(1, *ax_shape, dim).normal_(0, 1)))

    def forward(self, x):
        b, t, _ = x.shape
        embs = []
        for ind in range(len(self.shape)):
            ax_emb = getattr(self, f"weights_{ind}")
            expand_shape = (b, *self.shape, self.dim)
            embs.append(ax_emb.expand(expand_shape).reshape(b, self.max_seq_len, self.dim))
        pos_emb = sum(embs)
        return pos_emb[:, :t].to(x)


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_attr(o):
    if isinstance(o, dict):
        return _AttrDict({k: _to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_to_attr(v) for v in o]
    return o


def install():
    """Install shim modules into sys.modules and put the reference on sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if "axial_positional_embedding" not in sys.modules:
        m = types.ModuleType("axial_positional_embedding")
        m.AxialPositionalEmbedding = _AxialPositionalEmbedding
        sys.modules["axial_positional_embedding"] = m
    if "pytorch_lightning" not in sys.modules:
        m = types.ModuleType("pytorch_lightning")
        m.LightningModule = nn.Module
        sys.modules["pytorch_lightning"] = m
    if "omegaconf" not in sys.modules:
        import yaml
        m = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def load(path):
                with open(path) as f:
                    cfg = _to_attr(yaml.safe_load(f))
                cfg.model.params["lossconfig"] = _AttrDict(target="torch.nn.Identity")
                return cfg

        m.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = m
    import torchvision.io
    if not hasattr(torchvision.io, "write_video"):
        torchvision.io.write_video = lambda *a, **k: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


@contextlib.contextmanager
def ref_cwd():
    """vae.py:22 opens a path relative to the reference root."""
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def fake_clip_checkpoint(clip_state_dict):
    """Make `torch.jit.load(model_path)` (clip_model.py:535) return an object whose .state_dict()
    is `clip_state_dict` (CLIP-shaped, see weights.make_clip_state_dict)."""
    real = torch.jit.load

    class _Stub:
        def state_dict(self):
            return {k: v.clone() for k, v in clip_state_dict.items()}

    torch.jit.load = lambda *a, **k: _Stub()
    try:
        yield
    finally:
        torch.jit.load = real
