"""CPU oracle: a plain-PyTorch fp32 restatement of the MMVID token-generation hot path.

*** TEST INFRASTRUCTURE - NOT PRODUCT CODE. ***
Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs of `bench.py` may
import this file.  `mmvid_b200/` never imports it; the product path fails loudly without its CUDA library.

Parity pinning: the reference ships NO tests / golden vectors for this path.  The
oracle is therefore pinned against *the reference itself*, imported in the build container through
`oracle/ref_shims.py`: `tests/golden/gen_golden.py` runs reference modules and this restatement on the
same seeded weights/inputs, asserts they agree, and commits reference outputs as fixtures under
`tests/golden/` (`tests/test_oracle_vs_reference.py` re-checks live whenever /root/reference exists).
One boundary stays "parity unpinned": the third-party `axial_positional_embedding` package (pip,
unpinned, requirements.txt:2) is absent from /root/reference; both sides use the same restatement of its
published algorithm (see ref_shims.py).

Every function works from a *state dict with the reference's key names*  so the
same tensors can be fed to the reference, to this oracle and to the CUDA modules.  All citations are
relative to /root/reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# transformer  (mmvid_pytorch/transformers/clip_model.py)
# ------------------------------------------------------------------------------------------------


def build_attention_mask(context_length, mask_type="causal", index=()):
    """clip_model.py:561-578.  Additive float mask; 'mask_prev' rows i in `index` get -inf in [0, i)."""
    if mask_type == "causal":
        mask = torch.full((context_length, context_length), float("-inf")).triu_(1)
    elif mask_type == "mask_prev":
        mask = torch.zeros(context_length, context_length)
        for i in index:
            mask[i, :i] = float("-inf")
    else:
        raise NotImplementedError
    return mask


def quick_gelu(x):
    """clip_model.py:196-198."""
    return x * torch.sigmoid(1.702 * x)


def residual_attention_block(x, sd, p, n_head, attn_mask):
    """clip_model.py:201-227 on batch-first x [B,S,D] (the reference permutes to [S,B,D], :581-583;
    nn.MultiheadAttention math is batch-order independent).  LayerNorm eps 1e-5 (:188-193)."""
    B, S, D = x.shape
    hd = D // n_head
    h = F.layer_norm(x, (D,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
    qkv = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, n_head, hd).transpose(1, 2)
    k = k.view(B, S, n_head, hd).transpose(1, 2)
    v = v.view(B, S, n_head, hd).transpose(1, 2)
    att = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    if attn_mask is not None:
        att = att + attn_mask[:S, :S].to(att)
    att = torch.softmax(att, dim=-1)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, S, D)
    x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
    h = F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
    h = quick_gelu(h)
    x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    return x


def transformer_forward(x, sd, prefix, attn_mask, collect=None):
    """OpenAICLIPTransformer.forward (clip_model.py:580-584) -> Transformer (:230-247)."""
    n = 0
    D = x.shape[-1]
    n_head = D // 64  # build_model: heads = width // 64 (clip_model.py:466-504)
    while f"{prefix}resblocks.{n}.ln_1.weight" in sd:
        x = residual_attention_block(x, sd, f"{prefix}resblocks.{n}.", n_head, attn_mask)
        if collect is not None:
            collect.append(x)
        n += 1
    return x


# ------------------------------------------------------------------------------------------------
# axial positional embedding (third-party; see header) and the MMVID list wrapper (modules.py:8-53)
# ------------------------------------------------------------------------------------------------


def axial_pos_emb(sd, prefix, axial_shape, t, batch=1):
    D = sd[prefix + "weights_0"].shape[-1]
    total = int(np.prod(axial_shape))
    out = 0
    for i in range(len(axial_shape)):
        w = sd[prefix + f"weights_{i}"]
        out = out + w.expand(batch, *axial_shape, D).reshape(batch, total, D)
    return out[:, :t]


def axial_pos_emb_list(sd, prefix, num, axial_shape, batch=1):
    """modules.py:43-52 (no [SEP] branch)."""
    n = int(np.prod(axial_shape))
    return torch.cat([axial_pos_emb(sd, f"{prefix}module_list.{v}.", axial_shape, n, batch) for v in range(num)],
                     dim=1)


# ------------------------------------------------------------------------------------------------
# VQGAN  (taming/modules/diffusionmodules/model.py, taming/modules/vqvae/quantize.py, taming/models/vqgan.py)
# ------------------------------------------------------------------------------------------------


def _gn(x, sd, p):
    """Normalize: GroupNorm(32, C, eps=1e-6) (model.py:38-42)."""
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], 1e-6)


def _swish(x):
    """nonlinearity (model.py:33-35)."""
    return x * torch.sigmoid(x)


def _conv(x, sd, p, stride=1, padding=0):
    return F.conv2d(x, sd[p + "weight"], sd[p + "bias"], stride=stride, padding=padding)


def resnet_block(x, sd, p):
    """ResnetBlock.forward (model.py:130-150), temb=None, dropout 0."""
    h = _conv(_swish(_gn(x, sd, p + "norm1.")), sd, p + "conv1.", padding=1)
    h = _conv(_swish(_gn(h, sd, p + "norm2.")), sd, p + "conv2.", padding=1)
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(x, sd, p + "nin_shortcut.")
    return x + h


def attn_block(x, sd, p):
    """AttnBlock.forward (model.py:180-205): single-head spatial attention, scale C^-0.5."""
    h = _gn(x, sd, p + "norm.")
    q, k, v = _conv(h, sd, p + "q."), _conv(h, sd, p + "k."), _conv(h, sd, p + "v.")
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = torch.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, p + "proj_out.")


def _count(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}{n}.") for k in sd):
        n += 1
    return n


def vqgan_encoder(x, sd, p="encoder."):
    """Encoder.forward (model.py:439-466)."""
    n_levels = _count(sd, p + "down.")
    h = _conv(x, sd, p + "conv_in.", padding=1)
    for lvl in range(n_levels):
        nb = _count(sd, f"{p}down.{lvl}.block.")
        has_attn = _count(sd, f"{p}down.{lvl}.attn.") > 0
        for b in range(nb):
            h = resnet_block(h, sd, f"{p}down.{lvl}.block.{b}.")
            if has_attn:
                h = attn_block(h, sd, f"{p}down.{lvl}.attn.{b}.")
        if lvl != n_levels - 1:
            # Downsample (model.py:77-84): pad right/bottom by 1, 3x3 stride 2, no padding
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = _conv(h, sd, f"{p}down.{lvl}.downsample.conv.", stride=2)
    h = resnet_block(h, sd, p + "mid.block_1.")
    h = attn_block(h, sd, p + "mid.attn_1.")
    h = resnet_block(h, sd, p + "mid.block_2.")
    h = _conv(_swish(_gn(h, sd, p + "norm_out.")), sd, p + "conv_out.", padding=1)
    return h


def vqgan_decoder(z, sd, p="decoder."):
    """Decoder.forward (model.py:551-582)."""
    n_levels = _count(sd, p + "up.")
    h = _conv(z, sd, p + "conv_in.", padding=1)
    h = resnet_block(h, sd, p + "mid.block_1.")
    h = attn_block(h, sd, p + "mid.attn_1.")
    h = resnet_block(h, sd, p + "mid.block_2.")
    for lvl in reversed(range(n_levels)):
        nb = _count(sd, f"{p}up.{lvl}.block.")
        has_attn = _count(sd, f"{p}up.{lvl}.attn.") > 0
        for b in range(nb):
            h = resnet_block(h, sd, f"{p}up.{lvl}.block.{b}.")
            if has_attn:
                h = attn_block(h, sd, f"{p}up.{lvl}.attn.{b}.")
        if lvl != 0:
            # Upsample (model.py:56-62): nearest x2 then 3x3 conv
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"{p}up.{lvl}.upsample.conv.", padding=1)
    h = _conv(_swish(_gn(h, sd, p + "norm_out.")), sd, p + "conv_out.", padding=1)
    return h


def vq_distances(z_flat, codebook):
    """quantize.py:306-308: d = sum(z^2) + sum(e^2) - 2 z.e^T (this summation order)."""
    return torch.sum(z_flat ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) - \
        2 * torch.einsum("bd,dn->bn", z_flat, codebook.t())


def vq_indices(z, codebook):
    """VectorQuantizer2.forward (quantize.py:297-311): z [N,C,h,w] -> long [N*h*w] argmin indices."""
    z_flat = z.permute(0, 2, 3, 1).contiguous().view(-1, codebook.shape[1])
    return torch.argmin(vq_distances(z_flat, codebook), dim=1)


def vae_pre_quant(img, sd, p="model."):
    """encoder + quant_conv (vqgan.py:66-68) on img in [0,1] (vae.py:41: 2x-1)."""
    h = vqgan_encoder(2 * img - 1, sd, p + "encoder.")
    return _conv(h, sd, p + "quant_conv.")


def vae_get_codebook_indices(img, sd, p="model."):
    """VQGanVAE1024.get_codebook_indices (vae.py:38-43): img [N,3,H,W] in [0,1] -> long [N, n]."""
    z = vae_pre_quant(img, sd, p)
    idx = vq_indices(z, sd[p + "quantize.embedding.weight"])
    return idx.view(img.shape[0], -1)


def vae_decode(img_seq, sd, p="model."):
    """VQGanVAE1024.decode (vae.py:45-56): long [N,n] -> float [N,3,H,W] in [0,1]."""
    b, n = img_seq.shape
    z = F.embedding(img_seq, sd[p + "quantize.embedding.weight"])
    hw = int(math.sqrt(n))
    z = z.view(b, hw, hw, -1).permute(0, 3, 1, 2).contiguous()
    z = _conv(z, sd, p + "post_quant_conv.")
    img = vqgan_decoder(z, sd, p + "decoder.")
    return (img.clamp(-1.0, 1.0) + 1) * 0.5


def sub_state_dict(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------
# BERT  (mmvid_pytorch/dalle_bert.py)
# ------------------------------------------------------------------------------------------------


class BertSpec:
    """Shape bookkeeping of BERT.__init__ (dalle_bert.py:287-385)."""

    def __init__(self, *, dim, text_seq_len, num_text_tokens, num_visuals, num_targets, image_size,
                 num_image_tokens=1024, vae_layers=4, has_cvae=False, use_separate_visual_emb=False):
        self.dim = dim
        self.fmap = image_size // (2 ** vae_layers)
        self.image_seq_len = self.fmap ** 2
        self.num_visuals, self.num_targets = num_visuals, num_targets
        self.num_text_tokens = num_text_tokens + text_seq_len  # :303
        self.num_image_tokens = num_image_tokens
        self.text_seq_len = text_seq_len
        self.visual_seq_len = num_visuals * self.image_seq_len
        self.target_seq_len = num_targets * self.image_seq_len
        self.MASK = num_image_tokens       # :343-346
        self.SEP = num_image_tokens + 1
        self.separate_visual_emb = (has_cvae or use_separate_visual_emb) and num_visuals > 0  # :329-336
        self.rel_tok_index = 0
        self.st1_tok_index = 1 + text_seq_len + self.visual_seq_len  # :376
        self.vid_tok_index = self.st1_tok_index + 1                    # :377
        self.total_seq_len = 1 + text_seq_len + self.visual_seq_len + 2 + self.target_seq_len  # :380-385
        self.control_seq_len = 1 + text_seq_len + self.visual_seq_len + 2
        self.has_cvae = has_cvae

    def attn_mask(self):
        return build_attention_mask(self.total_seq_len, "mask_prev", [self.st1_tok_index, self.vid_tok_index])


def bert_control_emb(spec, sd, text, visual_tokens=None):
    """BERT.forward(return_loss=False) (dalle_bert.py:894-978) given visual *token ids* (long [B, V*n]) or None.

    (Raw visual frames go through cvae.get_codebook_indices first, :945; erase_* hooks are the caller's job.)
    """
    B = text.shape[0]
    dev = text.device
    before = torch.zeros(B, 1, dtype=torch.long, device=dev)  # [REL]=0 (:369, :903-906)
    control = F.embedding(before, sd["special_emb.weight"]) + F.embedding(before, sd["special_pos_emb.weight"])
    text_range = torch.arange(spec.text_seq_len, device=dev) + (spec.num_text_tokens - spec.text_seq_len)  # :917
    text = torch.where(text == 0, text_range, text)
    text_emb = F.embedding(text, sd["text_emb.weight"]) + sd["text_pos_emb.weight"][: spec.text_seq_len]
    control = torch.cat((control, text_emb), dim=1)
    if spec.num_visuals > 0:
        if visual_tokens is None:
            visual_tokens = torch.full((B, spec.visual_seq_len), spec.MASK, dtype=torch.long, device=dev)  # :955
        table = sd["visual_emb.weight"] if spec.separate_visual_emb else sd["image_emb.weight"]  # :959
        vemb = F.embedding(visual_tokens, table)
        vemb = vemb + axial_pos_emb_list(sd, "visual_pos_emb.", spec.num_visuals, (spec.fmap, spec.fmap), B)
        control = torch.cat((control, vemb), dim=1)
    after = torch.tensor([[1, 2]], dtype=torch.long, device=dev).repeat(B, 1)  # [ST1],[VID] (:370, :968-971)
    after_emb = F.embedding(after, sd["special_emb.weight"]) + F.embedding(after, sd["special_pos_emb.weight"])
    return torch.cat((control, after_emb), dim=1)


def erase_codebook_face(spec, visual_tokens, vc_mode, face_mode):
    """BERT.erase_codebook_face (dalle_bert.py:796-848) for the deterministic branches (face_mode given, or a vc_mode
    that never draws): visual-control token grids [B, V*n] with everything outside a window set to [MASK].
    The windows are hard-coded for 8 x 8 grids in the reference; they are applied as written whatever the fmap."""
    f = spec.fmap
    img = visual_tokens.view(visual_tokens.shape[0], -1, f, f)
    if vc_mode == "shape_4x4":
        out = img.clone()
        out[:, :, 1:3, 1:3] = spec.MASK  # :841
        return out.view_as(visual_tokens)
    out = torch.full_like(img, spec.MASK)
    if vc_mode == "face_8x8":
        assert face_mode is not None, "face_mode=None draws from random.random() (:806)"
        if face_mode == "eyes_nose":
            out[:, :, 2:5, 1:7] = img[:, :, 2:5, 1:7]   # :808
        else:
            out[:, :, 5:7, 2:6] = img[:, :, 5:7, 2:6]   # :810
    elif vc_mode == "face2_8x8":
        out[:, 0] = img[:, 0]                           # :815
        out[:, 1:, 2:6, 2:6] = img[:, 1:, 2:6, 2:6]     # :816
    elif vc_mode == "face3_8x8":
        out[:, 0] = img[:, 0]                           # :821
        out[:, :, 2:6, 2:6] = img[:, :, 2:6, 2:6]       # :822
    elif vc_mode in ("mask_8x8", "mask2_8x8"):
        assert face_mode is not None, "face_mode=None draws from np.random.choice (:826)"
        out[:, :, 1:7, 1:7] = img[:, :, 1:7, 1:7]       # strategy 3 (:829, :837-839)
    else:
        raise NotImplementedError(vc_mode)
    return out.view_as(visual_tokens)


def _to_logits(x, sd, p):
    """nn.Sequential(LayerNorm(dim), Linear) (dalle_bert.py:414-425)."""
    h = F.layer_norm(x, (x.shape[-1],), sd[p + "0.weight"], sd[p + "0.bias"], 1e-5)
    return F.linear(h, sd[p + "1.weight"], sd[p + "1.bias"])


def bert_target_pos_emb(spec, sd, batch=1):
    return axial_pos_emb(sd, "target_pos_emb.", (spec.num_targets, spec.fmap, spec.fmap), spec.target_seq_len, batch)


def bert_logits(spec, sd, control_emb, target_tokens, return_hidden=False):
    """Embed target ids (MASK allowed), run the transformer, image-token logits (dalle_bert.py:626-630)."""
    emb = F.embedding(target_tokens, sd["image_emb.weight"]) + bert_target_pos_emb(spec, sd, 1)
    tokens = torch.cat((control_emb, emb), dim=1)
    out = transformer_forward(tokens, sd, "transformer.transformer.", spec.attn_mask().to(tokens.device))
    logits = _to_logits(out[:, control_emb.shape[1]:], sd, "to_logits.")
    if return_hidden:
        return logits, out
    return logits


DEFAULT_MP_CONFIG = dict(  # utils/utils_args.py:221-281 defaults -> process_args :505-523
    T1_n=10, T2_n=10, T3_n=30, N1_n=0.9, N2_n=0.1, N3_n=0.125, N4_n=0.0625,
    T1_t=10, T2_t=5, T3_t=35, N1_t=0.0, N2_t=0.0, N3_t=0.0, N4_t=0.0, T=20, B=1)


def mask_predict_schedules(N, mp_config):
    """dalle_bert.py:594-614."""
    c = mp_config
    N3_n = max(1, int(N * c["N3_n"]))
    N4_n = max(1, int(N * c["N4_n"]))
    n = list(N * np.linspace(c["N1_n"], c["N2_n"], c["T1_n"])) + list(N3_n * np.ones(c["T2_n"])) + \
        list(N4_n * np.ones(c["T3_n"]))
    temp = list(np.linspace(c["N1_t"], c["N2_t"], c["T1_t"])) + list(c["N3_t"] * np.ones(c["T2_t"])) + \
        list(c["N4_t"] * np.ones(c["T3_t"]))
    return list(map(int, n)), temp


def _sample_multinomial(logits, temperature):
    """dalle_bert.py:527-538.  RNG order: rand_like(logits) then multinomial over [(b n), c]."""
    U = torch.rand_like(logits)
    g = -torch.log(-torch.log(U + 1e-20) + 1e-20)
    logits = logits + temperature * g
    probs = F.softmax(logits, dim=2)
    tok = torch.multinomial(probs.reshape(-1, probs.shape[-1]), 1).view(probs.shape[0], probs.shape[1], 1)
    Y = torch.gather(probs, 2, tok)
    return Y.squeeze(2), tok.squeeze(2)


@torch.no_grad()
def bert_mask_predict(spec, sd, control_emb, steps=10, mp_config=None, dynamic=True, preserve=None,
                      t_overlap=1, long_mode="long", trace=None):
    """BERT.mask_predict (dalle_bert.py:514-714).  Returns long [B, target_seq_len]."""
    mp_config = dict(DEFAULT_MP_CONFIG) if mp_config is None else mp_config
    dev = control_emb.device
    Ttot = spec.target_seq_len
    if long_mode == "long":
        if preserve is None:
            t_overlap = 0
        N = Ttot - spec.image_seq_len * t_overlap
    elif long_mode in ("interp", "interp2", "interp_real"):
        N = Ttot // 2
    else:
        N = Ttot
    preserve_mask1 = torch.zeros(1, Ttot, dtype=torch.long, device=dev)
    preserve_ = torch.full((control_emb.shape[0], Ttot), spec.MASK, dtype=torch.long, device=dev)
    if preserve is not None:
        if long_mode == "long":
            preserve_mask1[:, : spec.image_seq_len * t_overlap] = 1
            pr = preserve.reshape(-1, Ttot)  # '(b t) n -> b (t n)'
            preserve_[:, : spec.image_seq_len * t_overlap] = pr[:, -spec.image_seq_len * t_overlap:]
        else:
            pm = preserve_mask1.view(1, spec.num_targets, -1)
            pm[:, ::2, :] = 1
            pr = preserve.reshape(-1, spec.num_targets, spec.image_seq_len)
            pv = preserve_.view(-1, spec.num_targets, spec.image_seq_len)
            pv[:, ::2, :] = pr[:, : spec.num_targets // 2, :]
    no_preserve = preserve is None
    preserve = preserve_
    preserve_mask1 = preserve_mask1 == 1
    Tmax = mp_config["T"] if steps <= 0 else steps
    Bm = mp_config["B"]
    n, temp = mask_predict_schedules(N, mp_config)
    pos = bert_target_pos_emb(spec, sd, 1)
    mask_emb = sd["image_emb.weight"][spec.MASK]
    attn_mask = spec.attn_mask().to(dev)
    csl = control_emb.shape[1]

    def fwd(emb_target, c):
        tokens = torch.cat((c, emb_target + pos), dim=1)
        return transformer_forward(tokens, sd, "transformer.transformer.", attn_mask)

    samples = []
    for i in range(control_emb.shape[0]):
        c = control_emb[i:i + 1]
        tok_in = torch.full((1, Ttot), spec.MASK, dtype=torch.long, device=dev)
        if not no_preserve:
            tok_in[0] = torch.where(preserve_mask1[0], preserve[i], tok_in[0])
        out = fwd(F.embedding(tok_in, sd["image_emb.weight"]), c)
        logits = _to_logits(out[:, csl:], sd, "to_logits.")
        Y, I_new = _sample_multinomial(logits, temp[0])
        I_tok = torch.where(preserve_mask1, preserve[i:i + 1], I_new)
        if trace is not None:
            trace.append(dict(sample=i, t=0, tok=I_tok.clone(), logits=logits.clone()))
        Smax, tmax, Imax = 0, 0, None
        for t in range(1, Tmax):
            emb_in, masks1 = [], []
            for j in range(Bm):
                Y_valid = Y[~preserve_mask1]
                idx_valid = torch.arange(Ttot, device=dev)[~preserve_mask1[0]]
                try:
                    mask1_idx = torch.multinomial(Y_valid, N - n[t - 1], replacement=False)
                except RuntimeError:
                    mask1_idx = torch.multinomial(Y_valid, 1, replacement=False)
                mask1_idx = idx_valid[mask1_idx]
                mask1 = torch.zeros(Ttot, device=dev).scatter_(0, mask1_idx, 1).unsqueeze(0)
                mask1[preserve_mask1] = 1
                mask1 = mask1 == 1
                masks1.append(mask1)
                emb_out = F.embedding(I_tok, sd["image_emb.weight"])
                emb_in.append(torch.where(mask1.unsqueeze(2), emb_out, mask_emb))
            S = torch.zeros(Bm)
            YB, tokB = [], []
            for j in range(Bm):
                out = fwd(emb_in[j], c)
                logits = _to_logits(out[:, csl:], sd, "to_logits.")
                Y_new, I_new = _sample_multinomial(logits, temp[t])
                mask1_j = torch.bitwise_or(masks1[j], preserve_mask1)
                Y = torch.where(mask1_j, Y, Y_new)
                I_tok = torch.where(mask1_j, I_tok, I_new)
                s_rel = torch.sigmoid(_to_logits(out[:, spec.rel_tok_index], sd, "to_logits_rel."))
                s_vid = torch.sigmoid(_to_logits(out[:, spec.vid_tok_index], sd, "to_logits_vid."))
                S[j] = float(s_rel) * 0.5 + float(s_vid) * 0.5
                YB.append(Y)
                tokB.append(I_tok)
            jmax = int(S.argmax())
            Y, I_tok = YB[jmax], tokB[jmax]
            if trace is not None:
                trace.append(dict(sample=i, t=t, tok=I_tok.clone(), S=S.clone()))
            if dynamic:
                if S[jmax] > Smax:
                    tmax, Smax, Imax = t, S[jmax], I_tok
                if t - tmax >= 5:
                    break
            else:
                Imax = I_tok
        samples.append(Imax)
    return torch.cat(samples, 0)


@torch.no_grad()
def bert_generate_images(spec, sd, text, visual=None, steps=10, mp_config=None, dynamic=True, preserve=None,
                         t_overlap=1, long_mode="long"):
    """BERT.generate_images (dalle_bert.py:434-487) without the erase_visual / vc_mode hooks.
    `visual` = raw frames [B,V,3,H,W] (needs cvae.* or vae.* in sd) or None."""
    vis_tok = None
    if visual is not None and spec.num_visuals > 0:
        p = "cvae." if spec.has_cvae else "vae."
        b, v = visual.shape[:2]
        vis_tok = vae_get_codebook_indices(visual.reshape(b * v, *visual.shape[2:]), sub_state_dict(sd, p))
        vis_tok = vis_tok.view(b, -1)
    control = bert_control_emb(spec, sd, text, vis_tok)
    img_seq = bert_mask_predict(spec, sd, control, steps, mp_config, dynamic, preserve, t_overlap, long_mode)
    img_seq = img_seq.view(-1, spec.image_seq_len)
    images = vae_decode(img_seq, sub_state_dict(sd, "vae."))
    return images.view(text.shape[0], spec.num_targets, *images.shape[1:]), img_seq


# ------------------------------------------------------------------------------------------------
# ART-V  (mmvid_pytorch/dalle_artv.py)
# ------------------------------------------------------------------------------------------------


class ArtvSpec:
    """DALLE.__init__ bookkeeping (dalle_artv.py:122-187)."""

    def __init__(self, *, dim, text_seq_len, num_text_tokens, num_visuals, num_targets, image_size,
                 num_image_tokens=1024, vae_layers=4):
        assert num_visuals > 0
        self.dim = dim
        self.fmap = image_size // (2 ** vae_layers)
        self.image_seq_len = self.fmap ** 2
        self.num_visuals, self.num_targets = num_visuals, num_targets
        self.text_seq_len = text_seq_len
        self.target_seq_len = self.image_seq_len * num_targets
        self.visual_seq_len = self.image_seq_len * num_visuals
        self.control_seq_len = text_seq_len + self.visual_seq_len
        self.num_text_tokens = num_text_tokens + text_seq_len          # :132
        self.num_image_tokens = num_image_tokens
        self.num_visual_tokens = num_image_tokens + self.visual_seq_len  # :133
        self.num_control_tokens = self.num_text_tokens + self.num_visual_tokens  # :134
        self.total_tokens = self.num_text_tokens + num_image_tokens + self.num_visual_tokens  # :183
        self.total_seq_len = text_seq_len + self.target_seq_len + self.visual_seq_len  # :181

    def logits_mask(self):
        """dalle_artv.py:215-220 (True = forbidden)."""
        return (torch.block_diag(torch.ones(self.text_seq_len, self.num_text_tokens),
                                 torch.ones(self.visual_seq_len, self.num_visual_tokens),
                                 torch.ones(self.target_seq_len, self.num_image_tokens)) == 0).unsqueeze(0)


def artv_forward(spec, sd, text, visual_tokens=None, image_tokens=None):
    """DALLE.forward(return_loss=False) (dalle_artv.py:418-515) with visual already tokenised
    (long [B,V*n], -1 = erased) -> masked logits [B, seq, total_tokens]."""
    B, dev = text.shape[0], text.device
    text_range = torch.arange(spec.text_seq_len, device=dev) + (spec.num_text_tokens - spec.text_seq_len)
    text = torch.where(text == 0, text_range, text)
    text = F.pad(text, (1, 0), value=0)  # <bos> (:445-447)
    tokens = F.embedding(text, sd["text_emb.weight"]) + sd["text_pos_emb.weight"][: text.shape[1]]
    seq_len = text.shape[1]
    if visual_tokens is None:
        visual_tokens = -torch.ones(B, spec.visual_seq_len, dtype=torch.long, device=dev)  # :473
    visual_range = torch.arange(spec.visual_seq_len, device=dev) + (spec.num_visual_tokens - spec.visual_seq_len)
    visual_tokens = torch.where(visual_tokens == -1, visual_range, visual_tokens)  # :475-477
    vemb = F.embedding(visual_tokens, sd["visual_emb.weight"])
    vemb = vemb + axial_pos_emb_list(sd, "visual_pos_emb.", spec.num_visuals, (spec.fmap, spec.fmap), B)
    tokens = torch.cat((tokens, vemb), dim=1)
    seq_len += visual_tokens.shape[1]
    if image_tokens is not None and image_tokens.numel() > 0:
        shape = (spec.fmap, spec.fmap) if spec.num_targets == 1 else (spec.num_targets, spec.fmap, spec.fmap)
        iemb = F.embedding(image_tokens, sd["image_emb.weight"])
        iemb = iemb + axial_pos_emb(sd, "image_pos_emb.", shape, image_tokens.shape[1], B)
        tokens = torch.cat((tokens, iemb), dim=1)
        seq_len += image_tokens.shape[1]
    if tokens.shape[1] > spec.total_seq_len:  # :496-498
        seq_len -= 1
        tokens = tokens[:, :-1]
    mask = build_attention_mask(spec.total_seq_len, "causal").to(dev)
    out = transformer_forward(tokens, sd, "transformer.transformer.", mask)
    logits = _to_logits(out, sd, "to_logits.")
    lm = spec.logits_mask()[:, :seq_len].to(dev)
    return logits.masked_fill(lm, -torch.finfo(logits.dtype).max)  # :509-512


def top_k(logits, thres=0.5):
    """dalle_artv.py:61-67."""
    k = max(int((1 - thres) * logits.shape[-1]), 1)
    val, ind = torch.topk(logits, k)
    probs = torch.full_like(logits, float("-inf"))
    probs.scatter_(1, ind, val)
    return probs


@torch.no_grad()
def artv_generate_tokens(spec, sd, text, visual_tokens=None, filter_thres=0.5, temperature=1.0, max_new=None,
                         trace=None):
    """DALLE.generate_images sampling loop (dalle_artv.py:252-288): full re-forward per token (no cache)."""
    out = text[:, : spec.text_seq_len]
    total_len = spec.text_seq_len + spec.target_seq_len
    if max_new is not None:
        total_len = min(total_len, spec.text_seq_len + max_new)
    for cur_len in range(out.shape[1], total_len):
        txt, image = out[:, : spec.text_seq_len], out[:, spec.text_seq_len:]
        logits = artv_forward(spec, sd, txt, visual_tokens, image)[:, -1, :]
        if trace is not None:
            trace.append(logits.clone())
        probs = F.softmax(top_k(logits, filter_thres) / temperature, dim=-1)
        sample = torch.multinomial(probs, 1)
        sample = sample - spec.num_control_tokens  # is_image is always true here (:259, :278)
        out = torch.cat((out, sample), dim=-1)
    return out[:, spec.text_seq_len:]


# ------------------------------------------------------------------------------------------------
# BERT training losses  (mmvid_pytorch/dalle_bert.py:1030-1127) for GIVEN masks / negatives, so that the RNG-driven
# mask sampling (:992-1029) and frame warping (:204-238) stay outside the checked arithmetic.
# ------------------------------------------------------------------------------------------------


def bert_train_losses(spec, sd, text, visual_tokens, target_tokens, mask1, not_fully_masked, rel=False, vid=False,
                      rel_no_fully_masked=False, target_warp_tokens=None):
    """Returns (loss_msm, loss_rel, loss_vid); differentiable w.r.t. the tensors in `sd` that require grad.
    mask1: bool [B, T*n], True = ground-truth token kept (dalle_bert.py:1029-1031)."""
    B = text.shape[0]
    dev = text.device
    control = bert_control_emb(spec, sd, text, visual_tokens)
    csl = control.shape[1]
    attn_mask = spec.attn_mask().to(dev)
    pos = bert_target_pos_emb(spec, sd, 1)

    def fwd(c, tokens):
        emb = F.embedding(tokens, sd["image_emb.weight"]) + pos
        return transformer_forward(torch.cat((c, emb), dim=1), sd, "transformer.transformer.", attn_mask)

    tgt_masked = torch.where(mask1, target_tokens, torch.full_like(target_tokens, spec.MASK))
    out = fwd(control, tgt_masked)
    logits = _to_logits(out[:, csl:], sd, "to_logits.")
    loss_msm = F.cross_entropy(logits[~mask1], target_tokens[~mask1])  # :1040
    bce = F.binary_cross_entropy_with_logits
    denom = max(1.0, float(not_fully_masked.sum()))
    if rel:
        swapped = torch.cat(torch.chunk(control, 2, dim=0)[::-1], dim=0)  # swap() for an even batch (:110-113)
        out_neg = fwd(swapped, tgt_masked)
        lp = _to_logits(out[:, spec.rel_tok_index], sd, "to_logits_rel.").squeeze()
        ln = _to_logits(out_neg[:, spec.rel_tok_index], sd, "to_logits_rel.").squeeze()
        if rel_no_fully_masked:
            loss_rel = (bce(lp, torch.ones(B, device=dev), reduction="none") * not_fully_masked +
                        bce(ln, torch.zeros(B, device=dev), reduction="none") * not_fully_masked).sum() / denom
        else:
            loss_rel = bce(lp, torch.ones(B, device=dev)) + bce(ln, torch.zeros(B, device=dev))
    else:
        loss_rel = torch.tensor(0.0, device=dev)
    if vid and spec.num_targets > 1:
        warp_masked = torch.where(mask1, target_warp_tokens, torch.full_like(target_warp_tokens, spec.MASK))
        out_neg = fwd(control, warp_masked)
        lp = _to_logits(out[:, spec.vid_tok_index], sd, "to_logits_vid.")
        ln = _to_logits(out_neg[:, spec.vid_tok_index], sd, "to_logits_vid.")
        if rel_no_fully_masked:
            loss_vid = bce(lp, torch.ones(B, 1, device=dev), reduction="none").sum() / denom + \
                bce(ln, torch.zeros(B, 1, device=dev), reduction="none").sum() / denom
        else:
            loss_vid = bce(lp, torch.ones(B, 1, device=dev)) + bce(ln, torch.zeros(B, 1, device=dev))
    else:
        loss_vid = torch.tensor(0.0, device=dev)
    return loss_msm, loss_rel, loss_vid
