"""Deterministic synthetic weights and inputs (no checkpoint, no dataset, no network).

Released MMVID checkpoints (ViT-B-32.pt, vae_vox.ckpt, dalle.pt; reference README.md:29,141-151) cannot be
downloaded here, so every test / benchmark uses weights drawn by the recipes below.  The recipe depends
only on (key name, shape, seed) - not on module construction order - so the reference modules (in the
fixture generator), the CPU oracle and the CUDA modules can all be loaded with *identical* tensors through
`load_state_dict`.

Shapes / key names follow the reference state-dict contract.
"""
import hashlib
import math

import torch


def _gen(key, seed):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synth_tensor(key, shape, seed=0):
    """One tensor of the state dict, drawn from a per-key generator.

    Rules (chosen to look like the reference's default inits so activations stay O(1)):
      * norm / LayerNorm / GroupNorm `weight` (1-D, name contains 'norm' or 'ln_' or is `to_logits*.0`): 1 + 0.1 N(0,1)
      * any `bias`: 0.02 N(0,1)
      * embeddings (`*_emb.weight`, `weights_N`, `embedding.weight`): N(0,1) (nn.Embedding default)
      * everything else (Linear / Conv / in_proj): U(-b, b), b = 1/sqrt(fan_in)  (PyTorch kaiming_uniform(a=sqrt(5)))
    """
    g = _gen(key, seed)
    shape = tuple(shape)
    leaf = key.split(".")[-1]
    is_norm = (len(shape) == 1 and leaf == "weight")
    if is_norm:
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf in ("bias", "in_proj_bias"):
        return 0.02 * torch.randn(shape, generator=g)
    if leaf.startswith("weights_") or "emb.weight" in key or key.endswith("embedding.weight"):
        return torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    b = 1.0 / math.sqrt(max(fan_in, 1))
    return (torch.rand(shape, generator=g) * 2 - 1) * b


def fill_state_dict(module_or_sd, seed=0, fp16_round=(), codebook_std=None):
    """Return a new state dict with every floating tensor replaced by `synth_tensor`.

    fp16_round: key substrings whose tensors are rounded through fp16 (the reference's CLIP loader
      round-trips Linear / MultiheadAttention weights through half precision: clip_model.py:435-458,510,559).
    codebook_std: if given, `*.quantize.embedding.weight` ~ N(0, codebook_std^2) instead of N(0,1)
      (default VQ init U(+-1/1024), quantize.py:254, makes all distances nearly tie: SURVEY.md §7.2-4).
    """
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if not torch.is_floating_point(v):
            out[k] = v.clone()
            continue
        t = synth_tensor(k, v.shape, seed)
        if codebook_std is not None and k.endswith("quantize.embedding.weight"):
            t = t * codebook_std
        if any(s in k for s in fp16_round):
            t = t.half().float()
        out[k] = t.to(v.dtype)
    return out


# ---------------------------------------------------------------------------------------------
# CLIP-shaped checkpoint stand-in (what torch.jit.load('ViT-B-32.pt').state_dict() would hold)
# ---------------------------------------------------------------------------------------------

def resblock_keys(prefix, width, layers):
    ks = {}
    for n in range(layers):
        p = f"{prefix}resblocks.{n}."
        ks[p + "attn.in_proj_weight"] = (3 * width, width)
        ks[p + "attn.in_proj_bias"] = (3 * width,)
        ks[p + "attn.out_proj.weight"] = (width, width)
        ks[p + "attn.out_proj.bias"] = (width,)
        ks[p + "ln_1.weight"] = (width,)
        ks[p + "ln_1.bias"] = (width,)
        ks[p + "mlp.c_fc.weight"] = (4 * width, width)
        ks[p + "mlp.c_fc.bias"] = (4 * width,)
        ks[p + "mlp.c_proj.weight"] = (width, 4 * width)
        ks[p + "mlp.c_proj.bias"] = (width,)
        ks[p + "ln_2.weight"] = (width,)
        ks[p + "ln_2.bias"] = (width,)
    return ks


_FP16_KEYS = ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "c_fc.", "c_proj.")


def transformer_state_dict(width, layers, seed=0, prefix="transformer."):
    """Weights of one CLIP transformer stack with keys `<prefix>resblocks.N.*`.

    Values are fp16-representable where the reference's loader would have cast them to half
    (clip_model.py:435-458) so that `build_model(...).float()` reproduces them exactly.
    """
    sd = {}
    for k, shape in resblock_keys(prefix, width, layers).items():
        t = synth_tensor(k.replace(prefix, "clip."), shape, seed)
        if any(s in k for s in _FP16_KEYS):
            t = t.half().float()
        sd[k] = t
    return sd


def clip_checkpoint_state_dict(vision_width=768, vision_layers=12, text_width=512, text_layers=2, seed=0):
    """A minimal state dict that the reference's `build_model` (clip_model.py:461-512) accepts.

    Only `visual.transformer.*` (default --which_transformer openai_clip_visual, utils_args.py:183) or
    `transformer.*` is ever used by MMVID; every other tensor is a tiny placeholder.
    """
    patch, grid, embed, ctx, vocab = 32, 1, 8, 4, 8
    sd = {}
    sd["visual.conv1.weight"] = torch.zeros(vision_width, 3, patch, patch)
    sd["visual.class_embedding"] = torch.zeros(vision_width)
    sd["visual.positional_embedding"] = torch.zeros(grid * grid + 1, vision_width)
    sd["visual.ln_pre.weight"] = torch.ones(vision_width)
    sd["visual.ln_pre.bias"] = torch.zeros(vision_width)
    sd["visual.ln_post.weight"] = torch.ones(vision_width)
    sd["visual.ln_post.bias"] = torch.zeros(vision_width)
    sd["visual.proj"] = torch.zeros(vision_width, embed)
    vis = transformer_state_dict(vision_width, vision_layers, seed, prefix="transformer.")
    for k, v in vis.items():
        sd["visual." + k] = v
    sd["text_projection"] = torch.zeros(text_width, embed)
    sd["positional_embedding"] = torch.zeros(ctx, text_width)
    sd["token_embedding.weight"] = torch.zeros(vocab, text_width)
    sd["ln_final.weight"] = torch.ones(text_width)
    sd["ln_final.bias"] = torch.zeros(text_width)
    sd["logit_scale"] = torch.zeros(())
    txt = transformer_state_dict(text_width, text_layers, seed + 1, prefix="transformer.")
    sd.update(txt)
    sd["input_resolution"] = torch.tensor(patch * grid)
    sd["context_length"] = torch.tensor(ctx)
    sd["vocab_size"] = torch.tensor(vocab)
    return sd


# ---------------------------------------------------------------------------------------------
# Inputs (SURVEY.md §8d "Synthetic inputs")
# ---------------------------------------------------------------------------------------------

def synth_text(batch, text_seq_len, vocab=49408, seed=0, pad_frac=0.25):
    g = _gen("text", seed)
    t = torch.randint(1, vocab, (batch, text_seq_len), generator=g)
    npad = int(text_seq_len * pad_frac)
    if npad:
        t[:, text_seq_len - npad:] = 0
    return t


def synth_frames(batch, frames, image_size, seed=0):
    g = _gen("frames", seed)
    return torch.rand((batch, frames, 3, image_size, image_size), generator=g)


def synth_codes(batch, n, num_tokens=1024, seed=0):
    g = _gen("codes", seed)
    return torch.randint(0, num_tokens, (batch, n), generator=g)
