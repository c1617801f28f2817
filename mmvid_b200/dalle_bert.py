"""B200-native mirror of `mmvid_pytorch/dalle_bert.py:259-1127` (class BERT): the masked video-token model
with its mask-predict sampler.  Same constructor keywords, attributes, method signatures and state-dict keys
(`text_emb`, `text_pos_emb`, `image_emb`, `target_pos_emb.weights_*`, `visual_emb`, `visual_pos_emb.module_list.*`,
`special_emb`, `special_pos_emb`, `transformer.transformer.resblocks.*`, `to_logits*`, `vae.*`, `cvae.*`), so
`BERT(vae=..., cvae=..., **dalle_params)`, `load_state_dict(ckpt['weights'])`, `dalle(text, visual=..., return_loss=False)`
and `dalle.generate_images(...)` from the reference's train.py / test.py / utils_train.py work unchanged.

Sequence layout (dalle_bert.py:360-385):  [REL] text(L) visual(V*n) [ST1] [VID] target(T*n).

Compute: fused embedding gather -> transformer (mmvid_b200.transformer) -> LN+GEMM heads -> softmax, all in
libmmvid_b200.so.  torch is used for RNG (`rand_like` / `multinomial`, kept in PyTorch with the reference's
shapes and call order so sampled ids reproduce under a fixed seed), indexing glue and memory.

sampling_mode:
  'reference'  sample-serial, beam-serial loop with the reference's exact RNG consumption order (parity mode)
  'batched'    all samples advance together in one [B,S,D] forward per step; RNG order differs (documented),
               distributionally identical; used for throughput (mp_B == 1, dynamic=False only)
"""
import random

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import PRECISIONS
from .modules import AxialPositionalEmbedding, AxialPositionalEmbeddingList
from .transformer import OpenAICLIPTransformer


def exists(val):
    return val is not None


def set_requires_grad(model, value):
    if model is not None:
        for p in model.parameters():
            p.requires_grad = value


def eval_decorator(fn):
    def inner(model, *args, **kwargs):
        was_training = model.training
        model.eval()
        out = fn(model, *args, **kwargs)
        model.train(was_training)
        return out
    return inner


DEFAULT_MP_CONFIG = dict(T1_n=10, T2_n=10, T3_n=30, N1_n=0.9, N2_n=0.1, N3_n=0.125, N4_n=0.0625,
                         T1_t=10, T2_t=5, T3_t=35, N1_t=0.0, N2_t=0.0, N3_t=0.0, N4_t=0.0, T=20, B=1)


def mask_predict_schedules(N, mp_config):
    """Token re-mask counts n[t] and temperatures temp[t] (dalle_bert.py:594-614)."""
    c = mp_config
    n3 = max(1, int(N * c["N3_n"]))
    n4 = max(1, int(N * c["N4_n"]))
    n = list(N * np.linspace(c["N1_n"], c["N2_n"], c["T1_n"])) + [n3] * c["T2_n"] + [n4] * c["T3_n"]
    temp = list(np.linspace(c["N1_t"], c["N2_t"], c["T1_t"])) + [c["N3_t"]] * c["T2_t"] + [c["N4_t"]] * c["T3_t"]
    return [int(v) for v in n], [float(v) for v in temp]


class BERT(nn.Module):
    def __init__(self, *, dim, vae, cvae=None, num_text_tokens=10000, text_seq_len=256, stable=False,
                 text_feature_dim=0, fixed_language_model=None, which_transformer="none", num_visuals=1,
                 num_targets=1, use_separate_visual_emb=False, insert_sep=False, text_emb_bottleneck=False, **kwargs):
        super().__init__()
        if fixed_language_model is not None:
            raise NotImplementedError("fixed_language_model (RoBERTa features) is outside the hot path")
        if insert_sep:
            raise NotImplementedError("insert_sep layouts are outside the hot path")
        if stable:
            raise NotImplementedError("stable/DivideMax is never enabled by the reference CLI")
        image_size = vae.image_size
        num_image_tokens = vae.num_tokens
        fmap = vae.image_size // (2 ** vae.num_layers)
        self.dim = dim
        self.num_visuals, self.num_targets = num_visuals, num_targets

        num_text_tokens = num_text_tokens + text_seq_len  # unique pad id per position (dalle_bert.py:303)
        self.text_emb = nn.Embedding(num_text_tokens, dim)
        self.text_pos_emb = nn.Embedding(text_seq_len, dim)
        self.image_emb = nn.Embedding(num_image_tokens + 2, dim)
        self.target_pos_emb = AxialPositionalEmbedding(dim, axial_shape=(num_targets, fmap, fmap))
        if cvae is not None:
            use_separate_visual_emb = True
        if num_visuals > 0:
            self.visual_emb = nn.Embedding(num_image_tokens + 2, dim) if use_separate_visual_emb else None
            self.visual_pos_emb = AxialPositionalEmbeddingList(dim, num_visuals, axial_shape=(fmap, fmap))
        self.image_token_lut = {"[MASK]": num_image_tokens, "[SEP]": num_image_tokens + 1}

        self.num_text_tokens = num_text_tokens
        self.num_image_tokens = num_image_tokens
        self.text_seq_len = text_seq_len
        self.image_seq_len = fmap ** 2
        self.image_fmap_size = fmap
        self.image_size = image_size
        self.visual_seq_len = num_visuals * self.image_seq_len
        self.target_seq_len = num_targets * self.image_seq_len
        self.insert_sep = insert_sep

        self.special_token_lut = {"[REL]": 0, "[ST1]": 1, "[VID]": 2, "[ST3]": 3, "[ST4]": 4}
        self.num_special_tokens = len(self.special_token_lut)
        self.before_control_tok, self.after_control_tok = [0], [1, 2]
        self.before_control_seq_len, self.after_control_seq_len = 1, 2
        self.special_emb = nn.Embedding(self.num_special_tokens, dim)
        self.special_pos_emb = nn.Embedding(self.num_special_tokens, dim)
        self.rel_tok_index = 0
        self.st1_tok_index = 1 + self.text_seq_len + self.visual_seq_len
        self.vid_tok_index = self.st1_tok_index + 1
        self.txt_tok_index = 1
        self.total_seq_len = 1 + self.text_seq_len + self.visual_seq_len + 2 + self.target_seq_len

        self.vae, self.cvae = vae, cvae
        set_requires_grad(self.vae, False)
        set_requires_grad(self.cvae, False)

        self.fixed_language_model = None
        self.which_transformer = which_transformer
        if not which_transformer.startswith("openai_clip"):
            raise NotImplementedError
        self.transformer = OpenAICLIPTransformer(
            self.total_seq_len, which_transformer, model_path=kwargs.get("openai_clip_path"), causal=True,
            mask_type="mask_prev", mask_kwargs={"index": [self.st1_tok_index, self.vid_tok_index]},
            width=dim, layers=kwargs.get("transformer_layers"), precision=kwargs.get("precision", "tf32"))
        self.stable = False
        self.to_logits = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, self.num_image_tokens))
        self.to_logits_rel = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, 1))
        self.to_logits_vid = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, 1))
        self.current_step = 0
        self.sampling_mode = kwargs.get("sampling_mode", "reference")
        self.batch_train_passes = kwargs.get("batch_train_passes", True)  # positive + REL / VID negatives in one pass
        self._ids_cache = {}

    # ------------------------------------------------------------------------------------------ properties
    @property
    def precision(self):
        return self.transformer.precision

    @precision.setter
    def precision(self, p):
        assert p in PRECISIONS
        self.transformer.precision = p

    @property
    def control_seq_len(self):
        return 1 + self.text_seq_len + self.visual_seq_len + 2

    def _const_ids(self, vals, device):
        key = (tuple(vals), str(device))
        if key not in self._ids_cache:
            self._ids_cache[key] = torch.tensor([list(vals)], dtype=torch.long, device=device)
        return self._ids_cache[key]

    # ------------------------------------------------------------------------------------------ token helpers
    def get_image_tokens(self, image, reshape=True, insert_sep=False, which_vae="vae"):
        """dalle_bert.py:716-751: raw frames [B,T,3,H,W] (or list / 4-D) -> ids; id tensors pass through."""
        vae = self.cvae if (which_vae == "cvae" and self.cvae is not None) else self.vae
        if isinstance(image, list):
            assert len(image[0].shape) == 4, "image should be list of 4d image tensors"
            image = torch.stack(image, dim=1)
        if image.dim() == 4:
            image = image.unsqueeze(1)
        if image.dim() == 5:
            b, t, c, h, w = image.shape
            assert (c, h, w) == (3, vae.image_size, vae.image_size), \
                f"invalid image of dimensions {image.shape} passed in during training"
            ids = vae.get_codebook_indices(image.reshape(b * t, c, h, w))
            if insert_sep:
                raise NotImplementedError
            image = ids.view(b, t * ids.shape[1]) if reshape else ids
        return image

    @torch.no_grad()
    def recon_images(self, images, which_vae="vae"):
        vae = self.cvae if (which_vae == "cvae" and self.cvae is not None) else self.vae
        return vae.decode(self.get_image_tokens(images, reshape=False, which_vae=which_vae))

    @torch.no_grad()
    def get_codebook_emb(self, images, which_vae="vae"):
        b, t = images.shape[:2]
        img_seq = self.get_image_tokens(images, reshape=False, which_vae=which_vae)
        img_code = img_seq.view(b, t, -1)
        return img_code, F.embedding(img_code, self.image_emb.weight)

    def decode_images(self, img_seq):
        return self.vae.decode(img_seq.reshape(-1, self.image_seq_len))

    def decode_masks(self, mask):
        f = self.image_fmap_size
        patch = self.image_size // f
        m = mask.reshape(-1, 1, f, f)
        m = m.repeat_interleave(patch, 2).repeat_interleave(patch, 3)
        return F.pad(m, (0, 0, 0, 0, 0, 2))

    def random_erase_codebook(self, image, eraser, erase_half=False):
        """dalle_bert.py:779-794 (mutates `image` in place when erase_half, like the reference)."""
        f = self.image_fmap_size
        grid = image.view(image.shape[0], -1, f, f)
        if erase_half:
            grid[:, :, f // 2:, :] = self.image_token_lut["[MASK]"]
            out = grid
        else:
            out = torch.stack([eraser(c) for c in grid], dim=0)
        return out.reshape(image.shape[0], -1)

    # keep-windows (rows, cols) of the 8x8 token grid per visual-control mode (dalle_bert.py:796-848)
    _FACE_WINDOWS = {"eyes_nose": (slice(2, 5), slice(1, 7)), "mouth": (slice(5, 7), slice(2, 6))}

    def erase_codebook_face(self, image, vc_mode, face_mode=None):
        f = self.image_fmap_size
        MASK = self.image_token_lut["[MASK]"]
        grid = image.view(image.shape[0], -1, f, f)

        def keep_only(windows):
            out = torch.full_like(grid, MASK)
            for tsel, rs, cs in windows:
                out[:, tsel, rs, cs] = grid[:, tsel, rs, cs]
            return out

        every = slice(None)
        if vc_mode == "face_8x8":
            if face_mode is None:
                face_mode = "eyes_nose" if random.random() < 0.5 else "mouth"
            rs, cs = self._FACE_WINDOWS["eyes_nose" if face_mode == "eyes_nose" else "mouth"]
            grid = keep_only([(every, rs, cs)])
        elif vc_mode == "face2_8x8":
            grid = keep_only([(slice(0, 1), every, every), (slice(1, None), slice(2, 6), slice(2, 6))])
        elif vc_mode == "face3_8x8":
            grid = keep_only([(slice(0, 1), every, every), (every, slice(2, 6), slice(2, 6))])
        elif vc_mode in ("mask_8x8", "mask2_8x8"):
            strategy = int(np.random.choice([1, 2, 3], p=[0.5, 0.25, 0.25])) if face_mode is None else 3
            if strategy == 2:
                grid = keep_only([(every, slice(2, 6), slice(2, 6))])
            elif strategy == 3:
                grid = keep_only([(every, slice(1, 7), slice(1, 7))])
        elif vc_mode == "shape_4x4":
            grid[:, :, 1:3, 1:3] = MASK
        else:
            raise NotImplementedError
        return grid.reshape(image.shape[0], -1)

    def get_special_token(self, tok_list, batch_size=1, device="cuda"):
        return torch.tensor(tok_list, dtype=torch.long, device=device).repeat(batch_size, 1)

    # ------------------------------------------------------------------------------------------ embeddings
    def _control_segments(self, text, visual_ids):
        dev = text.device
        segs = [dict(ids=self._const_ids([0], dev), seq_off=0, table=self.special_emb.weight.detach(),
                     table2=self.special_pos_emb.weight.detach()),
                dict(ids=text.contiguous(), seq_off=1, table=self.text_emb.weight.detach(),
                     pos=self.text_pos_emb.weight.detach(),
                     pad=(0, self.num_text_tokens - self.text_seq_len))]
        off = 1 + self.text_seq_len
        if self.num_visuals > 0:
            table = self.visual_emb.weight if self.visual_emb is not None else self.image_emb.weight
            segs.append(dict(ids=visual_ids.contiguous(), seq_off=off, table=table.detach(),
                             pos=self.visual_pos_emb.table()))
            off += self.visual_seq_len
        segs.append(dict(ids=self._const_ids([1, 2], dev), seq_off=off, table=self.special_emb.weight.detach(),
                         table2=self.special_pos_emb.weight.detach()))
        return segs

    def _target_segment(self, ids):
        return dict(ids=ids.contiguous(), seq_off=self.control_seq_len, table=self.image_emb.weight.detach(),
                    pos=self.target_pos_emb.table())

    def transformer_forward(self, tokens):
        return self.transformer(tokens)

    def _head(self, rows, seq):
        """nn.Sequential(LayerNorm, Linear) on [M, D] rows (dalle_bert.py:414-425)."""
        prec = PRECISIONS[self.precision]
        h = ops.layernorm(rows, seq[0].weight, seq[0].bias, 1e-5, out_dtype=ops.act_dtype(prec))
        w = self.transformer._w(seq[1].weight, prec)
        return ops.linear(h, w, seq[1].bias, precision=prec)

    def _head_scalar(self, row, seq):
        """to_logits_rel / to_logits_vid on a [M<=16, D] slice (fp32 GEMV)."""
        h = ops.layernorm(row, seq[0].weight, seq[0].bias, 1e-5)
        return ops.linear_small_m(h, seq[1].weight.detach(), seq[1].bias)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, text, visual=None, target=None, mask=None, return_loss=False, rel=False, vid=False,
                erase_visual=False, erase_visual_half=False, msm_strategy_prob=[0.7, 0.1, 0.1, 0.1],
                msm_bernoulli_prob=[0.2, 0.5], rel_no_fully_masked=False, vid_strategy_prob=[0.25, 0.25, 0.25, 0.25],
                negvc=False, visual_neg=None, text_neg=None, pc_prob=0, vc_mode=None, face_mode=None,
                visual_aug_mode=None, **kwargs):
        """return_loss=False: control embedding [B, 1+L+V*n+2, D] (dalle_bert.py:894-978)."""
        assert text.shape[-1] == self.text_seq_len, \
            f"the length {text.shape[-1]} of the text tokens you passed in does not have the correct length ({self.text_seq_len})"
        B, dev = text.shape[0], text.device
        visual_ids = None
        if self.num_visuals > 0:
            if exists(visual) and len(visual):
                if visual_aug_mode is not None:
                    from .augment import augment_visual
                    visual = augment_visual(visual, visual_aug_mode)  # dalle_bert.py:940-944
                visual_ids = self.get_image_tokens(visual, insert_sep=self.insert_sep, which_vae="cvae")
                if erase_visual:
                    visual_ids = self.random_erase_codebook(visual_ids, self._visual_eraser(), erase_visual_half)
                if vc_mode is not None:
                    visual_ids = self.erase_codebook_face(visual_ids, vc_mode, face_mode)
            else:
                visual_ids = torch.full((B, self.visual_seq_len), self.image_token_lut["[MASK]"], dtype=torch.long,
                                        device=dev)
        with torch.no_grad():
            control = torch.empty(B, self.control_seq_len, self.dim, device=dev, dtype=torch.float32)
            ops.embed_gather(control, self._control_segments(text, visual_ids))
        if not return_loss:
            return control
        # ------------------------------------------------------------------ training losses (dalle_bert.py:980-1127)
        text_neg_ids = None
        if negvc:
            # dalle_bert.py:909-935, 974-975: the negative control sequence is [REL] + text_neg + [ST1][VID]; the reference
            # never adds a visual segment to it (visual_neg is accepted and ignored), so with visual control it is shorter
            # than the positive one and the mask rows keep their absolute positions (clip_model.py:217-221 slices the mask).
            assert text_neg is not None and text_neg.shape == text.shape, "negvc=True needs text_neg shaped like text"
            text_neg_ids = text_neg
        target_orig = target
        target_ids = self.get_image_tokens(target)
        mask1, not_fully_masked = self._sample_msm_masks(B, dev, msm_strategy_prob, msm_bernoulli_prob, pc_prob)
        target_warp_ids = None
        if vid and self.num_targets > 1:
            from .augment import warp
            target_warp_ids = self.get_image_tokens(warp(target_orig, vid_strategy_prob))
        return self._losses(text, visual_ids, target_ids, mask1, not_fully_masked, rel=rel, vid=vid,
                            rel_no_fully_masked=rel_no_fully_masked, target_warp_ids=target_warp_ids,
                            text_neg=text_neg_ids)

    # ------------------------------------------------------------------------------------------ training internals
    def _sample_msm_masks(self, B, dev, strategy_prob, bernoulli_prob, pc_prob):
        """Per-sample keep-masks for masked sequence modelling (dalle_bert.py:992-1029): True = ground truth kept.
        Strategies: 1 bernoulli keep with p~U(a,b); 2 mask everything; 3 keep outside a random box; 4 keep inside it."""
        import torchvision.transforms as T
        if not hasattr(self, "random_erasing"):
            self.random_erasing = T.RandomErasing(p=1, scale=(0.2, 0.8), ratio=(0.5, 2), value=0)
        f, Tn, n = self.image_fmap_size, self.num_targets, self.image_seq_len
        masks, nfm = [], torch.ones(B, device=dev)
        for i in range(B):
            strategy = int(np.random.choice([1, 2, 3, 4], p=strategy_prob))
            if strategy == 1:
                p = np.random.uniform(*bernoulli_prob)
                m = torch.bernoulli(torch.ones(self.target_seq_len, device=dev) * p)
            elif strategy == 2:
                nfm[i] = 0
                m = torch.zeros(self.target_seq_len, device=dev)
            else:
                box = self.random_erasing(torch.ones(Tn, 1, f, f, device=dev)).reshape(-1)
                m = box if strategy == 3 else 1 - box
            if pc_prob > 0 and random.random() < pc_prob:
                for tt in random.sample(range(Tn), random.randint(1, Tn // 2)):
                    m[n * tt:n * (tt + 1)] = 1
            masks.append(m)
        return torch.stack(masks, 0) == 1, nfm

    def _embed_train(self, text, visual_ids, target_ids_masked, with_visual=True, with_target=True):
        """Residual stream [B,S,D] with autograd to every embedding table (same fused gather kernel as inference)."""
        from .autograd import EmbedFn
        B, dev, D = text.shape[0], text.device, self.dim

        def seg(table, table2, pos, ids, pad=None):
            n = ids.shape[1]
            buf = torch.empty(B, n, D, device=dev, dtype=torch.float32)
            return EmbedFn.apply(table, table2, pos, ids.contiguous(), B, n, 0, pad, buf)

        parts = [seg(self.special_emb.weight, self.special_pos_emb.weight, None, self._const_ids([0], dev).expand(B, 1).contiguous()),
                 seg(self.text_emb.weight, None, self.text_pos_emb.weight, text, (0, self.num_text_tokens - self.text_seq_len))]
        if self.num_visuals > 0 and with_visual:
            table = self.visual_emb.weight if self.visual_emb is not None else self.image_emb.weight
            parts.append(seg(table, None, self._axial_pos(self.visual_pos_emb), visual_ids))
        parts.append(seg(self.special_emb.weight, self.special_pos_emb.weight, None,
                         self._const_ids([1, 2], dev).expand(B, 2).contiguous()))
        control = torch.cat(parts, dim=1)
        if not with_target:
            return control, None
        target = seg(self.image_emb.weight, None, self._axial_pos(self.target_pos_emb), target_ids_masked)
        return control, target

    @staticmethod
    def _axial_pos(mod):
        """Differentiable [n, D] position table (sum of broadcast axis tables) for the training path."""
        mods = list(mod.module_list) if hasattr(mod, "module_list") else [mod]
        tabs = []
        for m in mods:
            ws = m.axis_weights()
            t = ws[0]
            for w in ws[1:]:
                t = t + w
            tabs.append(t.reshape(-1, m.dim))
        return torch.cat(tabs, 0) if len(tabs) > 1 else tabs[0]

    def _transformer_train(self, tokens):
        from .autograd import AttentionFn, LayerNormFn, LinearFn
        from ._lib import ACT_NONE, ACT_QUICKGELU, H16
        tr = self.transformer
        prec = PRECISIONS[tr.precision]
        if prec in H16:
            prec = 1  # training runs the tf32 tensor-core path (16-bit activations are an inference-only mode)
        B, S, D = tokens.shape
        H = tr.transformer.heads
        x = tokens.reshape(B * S, D)
        for blk in tr.transformer.resblocks:
            h = LayerNormFn.apply(x, blk.ln_1.weight, blk.ln_1.bias, 1e-5)
            qkv = LinearFn.apply(h, blk.attn.in_proj_weight, blk.attn.in_proj_bias, ACT_NONE, prec)
            att = AttentionFn.apply(qkv, B, S, H, tr.mask_kind, tr.mask_rows, prec)
            x = x + LinearFn.apply(att, blk.attn.out_proj.weight, blk.attn.out_proj.bias, ACT_NONE, prec)
            h = LayerNormFn.apply(x, blk.ln_2.weight, blk.ln_2.bias, 1e-5)
            h = LinearFn.apply(h, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, ACT_QUICKGELU, prec)
            x = x + LinearFn.apply(h, blk.mlp.c_proj.weight, blk.mlp.c_proj.bias, ACT_NONE, prec)
        return x.view(B, S, D)

    def _head_train(self, rows, seq):
        from .autograd import LayerNormFn, LinearFn
        from ._lib import ACT_NONE, FP32
        h = LayerNormFn.apply(rows, seq[0].weight, seq[0].bias, 1e-5)
        prec = PRECISIONS[self.transformer.precision]
        if prec in (2, 3):
            prec = 1                      # bf16 / fp16 are inference-only; training GEMMs run tf32
        if seq[1].weight.shape[0] < 8:
            prec = FP32                   # scalar REL / VID heads: CUDA-core path
        return LinearFn.apply(h, seq[1].weight, seq[1].bias, ACT_NONE, prec)

    def _losses(self, text, visual_ids, target_ids, mask1, not_fully_masked, rel=False, vid=False,
                rel_no_fully_masked=False, target_warp_ids=None, swap_perm=None, text_neg=None):
        """MSM / REL / VID losses (dalle_bert.py:1030-1127) for given masks; autograd flows to all trainable params."""
        from .autograd import cross_entropy_selected
        B, dev, D = text.shape[0], text.device, self.dim
        MASK = self.image_token_lut["[MASK]"]
        csl, Ttot = self.control_seq_len, self.target_seq_len
        tgt_masked = torch.where(mask1, target_ids, torch.full_like(target_ids, MASK))
        control, target_emb = self._embed_train(text, visual_ids, tgt_masked)
        # The positive pass and the REL / VID negatives share the transformer and (unless negvc shortens the control
        # sequence) the sequence length: they run as ONE pass over a 2B / 3B batch (the reference runs
        # three passes, dalle_bert.py:1037,1057,1101).  Samples are independent inside the transformer, so the losses
        # and gradients are those of the separate passes.
        passes = [torch.cat((control, target_emb), dim=1)]
        i_rel = i_vid = None
        out_neg_rel = None
        if rel:
            assert B >= 2 and B % 2 == 0, "REL needs an even batch (control sequences are swapped between halves)"
            if text_neg is not None:  # negvc: explicit negative text, no visual segment (dalle_bert.py:1047-1055)
                control_neg = self._embed_train(text_neg, None, tgt_masked, with_visual=False, with_target=False)[0]
            else:
                perm = swap_perm if swap_perm is not None else torch.cat((torch.arange(B // 2, B), torch.arange(0, B // 2))).to(dev)
                control_neg = control[perm]
            neg = torch.cat((control_neg, target_emb), dim=1)
            if neg.shape[1] == passes[0].shape[1] and self.batch_train_passes:
                i_rel = len(passes)
                passes.append(neg)
            else:
                out_neg_rel = self._transformer_train(neg)
        do_vid = vid and self.num_targets > 1
        if do_vid:
            warp_masked = torch.where(mask1, target_warp_ids, torch.full_like(target_warp_ids, MASK))
            _, warp_emb = self._embed_train(text, visual_ids, warp_masked)
            neg = torch.cat((control, warp_emb), dim=1)
            if self.batch_train_passes:
                i_vid = len(passes)
                passes.append(neg)
            else:
                out_neg_vid = self._transformer_train(neg)
        outs = self._transformer_train(torch.cat(passes, dim=0) if len(passes) > 1 else passes[0])
        out = outs[:B]
        if i_rel is not None:
            out_neg_rel = outs[i_rel * B:(i_rel + 1) * B]
        if i_vid is not None:
            out_neg_vid = outs[i_vid * B:(i_vid + 1) * B]
        logits = self._head_train(out[:, csl:].reshape(B * Ttot, D), self.to_logits)
        loss_msm = cross_entropy_selected(logits, target_ids.reshape(-1), (~mask1).reshape(-1))
        bce = F.binary_cross_entropy_with_logits
        ones, zeros = torch.ones(B, device=dev), torch.zeros(B, device=dev)
        denom = max(1.0, float(not_fully_masked.sum()))
        if rel:
            lp = self._head_train(out[:, self.rel_tok_index], self.to_logits_rel).squeeze(-1)
            ln = self._head_train(out_neg_rel[:, self.rel_tok_index], self.to_logits_rel).squeeze(-1)
            if rel_no_fully_masked:
                loss_rel = ((bce(lp, ones, reduction="none") + bce(ln, zeros, reduction="none")) * not_fully_masked).sum() / denom
            else:
                loss_rel = bce(lp, ones) + bce(ln, zeros)
        else:
            loss_rel = torch.tensor(0.0, device=dev)
        if do_vid:
            lp = self._head_train(out[:, self.vid_tok_index], self.to_logits_vid)
            ln = self._head_train(out_neg_vid[:, self.vid_tok_index], self.to_logits_vid)
            o1, z1 = torch.ones(B, 1, device=dev), torch.zeros(B, 1, device=dev)
            if rel_no_fully_masked:
                loss_vid = bce(lp, o1, reduction="none").sum() / denom + bce(ln, z1, reduction="none").sum() / denom
            else:
                loss_vid = bce(lp, o1) + bce(ln, z1)
        else:
            loss_vid = torch.tensor(0.0, device=dev)
        return loss_msm, loss_rel, loss_vid

    def _visual_eraser(self):
        import torchvision.transforms as T
        if not hasattr(self, "visual_eraser"):
            self.visual_eraser = T.RandomErasing(p=0.95, scale=(0.55, 0.85), ratio=(0.5, 2), value=self.num_image_tokens)
        return self.visual_eraser

    # ------------------------------------------------------------------------------------------ sampling
    @torch.no_grad()
    @eval_decorator
    def generate_images(self, text, *, visual=None, mask=None, img=None, argmax=False, dynamic=True, debug=False,
                        erase_visual=False, mask_predict_steps=10, preserve=None, t_overlap=1, pc_mode=None,
                        vc_mode=None, face_mode=None, mp_config=None, long_mode="long"):
        """dalle_bert.py:434-487 -> (images [b,t,3,H,W], pnag_samples, img_seq [(b t), n])."""
        control_emb = self(text, visual=visual, erase_visual=erase_visual, erase_visual_half=True, vc_mode=vc_mode,
                           face_mode=face_mode, return_loss=False)
        img_seq, pnag_samples = self.mask_predict(control_emb, argmax=argmax, dynamic=dynamic, debug=debug,
                                                  steps=mask_predict_steps, preserve=preserve, t_overlap=t_overlap,
                                                  pc_mode=pc_mode, mp_config=mp_config, long_mode=long_mode)
        img_seq = img_seq.reshape(-1, self.image_seq_len)
        images = self.vae.decode(img_seq)
        images = images.view(-1, self.num_targets, *images.shape[1:])
        return images, pnag_samples, img_seq

    def _sample_multinomial(self, logits, temperature):
        """dalle_bert.py:527-538 with identical RNG consumption: rand_like(logits) then multinomial([(b n), c], 1)."""
        U = torch.rand_like(logits)
        noise = None
        if temperature != 0:
            noise = (-torch.log(-torch.log(U + 1e-20) + 1e-20)).contiguous()
        probs = ops.softmax_logits(logits, noise, float(temperature))
        tok = torch.multinomial(probs.view(-1, probs.shape[-1]), 1).view(probs.shape[0], probs.shape[1], 1)
        Y = torch.gather(probs, 2, tok)
        return Y.squeeze(2), tok.squeeze(2)

    def _preserve_setup(self, nb, preserve, t_overlap, long_mode, dev):
        """dalle_bert.py:541-583."""
        Ttot, n = self.target_seq_len, self.image_seq_len
        MASK = self.image_token_lut["[MASK]"]
        if long_mode == "long":
            if preserve is None:
                t_overlap = 0
            N = Ttot - n * t_overlap
        elif long_mode in ("interp", "interp2", "interp_real"):
            N = Ttot // 2
        else:
            N = Ttot
        pmask = torch.zeros(1, Ttot, dtype=torch.bool, device=dev)
        ptok = torch.full((nb, Ttot), MASK, dtype=torch.long, device=dev)
        if preserve is not None:
            if long_mode == "long":
                pmask[:, : n * t_overlap] = True
                pr = preserve.reshape(-1, Ttot)
                ptok[:, : n * t_overlap] = pr[:, Ttot - n * t_overlap:]
            elif long_mode in ("interp", "interp2", "interp_real"):
                pmask.view(1, self.num_targets, n)[:, ::2, :] = True
                pr = preserve.reshape(-1, self.num_targets, n)
                ptok.view(nb, self.num_targets, n)[:, ::2, :] = pr[:, : self.num_targets // 2, :]
        return N, pmask, ptok

    @torch.no_grad()
    def mask_predict(self, control_emb, dynamic=True, debug=False, steps=10, preserve=None, t_overlap=1,
                     mp_config=None, long_mode="long", **kwargs):
        """dalle_bert.py:514-714 -> (long [B, target_seq_len], image_samples)."""
        mp_config = dict(DEFAULT_MP_CONFIG) if mp_config is None else mp_config
        if self.sampling_mode == "batched" and not debug:
            import os
            if os.environ.get("MMVID_SAMPLER", "device") == "torch" and mp_config["B"] == 1 and not dynamic:
                return self._mask_predict_batched(control_emb, steps, preserve, t_overlap, mp_config, long_mode), []
            return self._mask_predict_device(control_emb, steps, preserve, t_overlap, mp_config, long_mode, dynamic), []
        dev = control_emb.device
        nb, csl, D = control_emb.shape
        Ttot = self.target_seq_len
        MASK = self.image_token_lut["[MASK]"]
        N, pmask, ptok = self._preserve_setup(nb, preserve, t_overlap, long_mode, dev)
        Tmax = mp_config["T"] if steps <= 0 else steps
        Bm = mp_config["B"]
        n, temp = mask_predict_schedules(N, mp_config)
        arange = torch.arange(Ttot, device=dev)
        x = torch.empty(1, self.total_seq_len, D, device=dev, dtype=torch.float32)
        image_samples, sample_toks = [], []

        def run(ids_in):
            ops.embed_gather(x, [self._target_segment(ids_in)])
            out = self.transformer(x)
            return out, self._head(out[0, csl:], self.to_logits).view(1, Ttot, -1)

        for i in range(nb):
            x[:, :csl].copy_(control_emb[i:i + 1])
            tok_in = torch.where(pmask, ptok[i:i + 1], torch.full_like(ptok[i:i + 1], MASK))
            out, logits = run(tok_in)
            Y, I_new = self._sample_multinomial(logits, temp[0])
            I_tok = torch.where(pmask, ptok[i:i + 1], I_new)
            if debug:
                image_samples.append(self.decode_images(I_tok))
            Smax, tmax, Imax = 0, 0, None
            for t in range(1, Tmax):
                keeps, ids_in = [], []
                for j in range(Bm):
                    Y_valid = Y[~pmask]
                    idx_valid = arange[~pmask[0]]
                    try:
                        kept = torch.multinomial(Y_valid, N - n[t - 1], replacement=False)
                    except RuntimeError:
                        kept = torch.multinomial(Y_valid, 1, replacement=False)
                    keep = torch.zeros(Ttot, dtype=torch.bool, device=dev)
                    keep[idx_valid[kept]] = True
                    keep = (keep | pmask[0]).unsqueeze(0)
                    keeps.append(keep)
                    ids_in.append(torch.where(keep, I_tok, torch.full_like(I_tok, MASK)))
                S = torch.zeros(Bm)
                YB, tokB = [], []
                for j in range(Bm):
                    out, logits = run(ids_in[j])
                    Y_new, I_new = self._sample_multinomial(logits, temp[t])
                    Y = torch.where(keeps[j], Y, Y_new)
                    I_tok = torch.where(keeps[j], I_tok, I_new)
                    if dynamic or Bm > 1:
                        rows = out[0, [self.rel_tok_index, self.vid_tok_index]]
                        s_rel = torch.sigmoid(self._head_scalar(rows[0:1].contiguous(), self.to_logits_rel))
                        s_vid = torch.sigmoid(self._head_scalar(rows[1:2].contiguous(), self.to_logits_vid))
                        S[j] = float(s_rel) * 0.5 + float(s_vid) * 0.5
                    YB.append(Y)
                    tokB.append(I_tok)
                jmax = int(S.argmax())
                Y, I_tok = YB[jmax], tokB[jmax]
                if debug:
                    mask_img = self.decode_masks((~keeps[jmax]).float())
                    masked_img = torch.clamp(image_samples[-1] * 0.7 + mask_img * 0.4, 0, 1)
                    image_samples.append(masked_img)
                    image_samples.append(self.decode_images(I_tok))
                if dynamic:
                    if S[jmax] > Smax:
                        tmax, Smax, Imax = t, S[jmax], I_tok
                    if t - tmax >= 5:
                        break
                else:
                    Imax = I_tok
            sample_toks.append(Imax)
        return torch.cat(sample_toks, 0), image_samples

    # ------------------------------------------------------------------------------------------ CUDA graph of one forward
    def _use_cuda_graph(self, dev):
        import os
        return dev.type == "cuda" and os.environ.get("MMVID_CUDA_GRAPH", "1") != "0" and not torch.is_grad_enabled() \
            and not torch.cuda.is_current_stream_capturing()

    def _weights_version(self):
        """Changes whenever a parameter is updated in place or replaced (training step, load_state_dict, .half())."""
        v = 0
        for p in self.parameters():
            v = (v * 1000003 + p._version * 31 + p.data_ptr()) & 0xFFFFFFFFFFFF
        from . import _lib
        return (v, _lib.weights_epoch())

    def _forward_graph(self, nb, dev):
        """dict(graph, x [nb,S,D], ids [nb,Ttot], logits [nb,Ttot,1024]): target-embedding gather -> transformer ->
        logits head captured in one CUDA graph.  Static buffers: the caller fills x[:, :control] and ids, replays,
        and reads logits before the next replay (same stream)."""
        key = (nb, str(dev), str(self.precision), str(self.transformer.precision), self._weights_version())
        cache = getattr(self, "_graph_cache", None)
        if cache is None or not isinstance(cache, dict) or cache.get("_sig") != key[1:]:
            cache = self._graph_cache = {"_sig": key[1:]}  # new weights / precision / device: drop every captured graph
        ent = cache.get(nb)
        if ent is not None:
            return ent
        while len(cache) > 4:  # "_sig" + at most 3 batch sizes (first step, beams, a second caller)
            cache.pop(next(k for k in cache if k != "_sig"))
        D, Ttot, csl = self.dim, self.target_seq_len, self.control_seq_len
        x = torch.zeros(nb, self.total_seq_len, D, device=dev, dtype=torch.float32)
        ids = torch.full((nb, Ttot), self.image_token_lut["[MASK]"], dtype=torch.long, device=dev)

        score_idx = torch.tensor([self.rel_tok_index, self.vid_tok_index], device=dev)

        def fwd():
            ops.embed_gather(x, [self._target_segment(ids)])
            out = self.transformer(x)
            # [REL] / [VID] rows for the beam / dynamic-stop scores (dalle_bert.py:685-689)
            return (self._head(out[:, csl:].reshape(nb * Ttot, D), self.to_logits).view(nb, Ttot, -1),
                    out.index_select(1, score_idx).contiguous())

        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up: one-time kernel attributes, bf16 weight copies, Q/K/V buffers
            fwd()
            fwd()
        cur.wait_stream(side)
        from . import _lib
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            logits, score_rows = fwd()
        # The graph holds raw pointers: keep every tensor it reads or writes alive for as long as the graph lives, even if
        # a later forward of another shape makes the transformer replace its cached Q / K / V^T buffers or bf16 copies.
        keep = [getattr(self.transformer, "_qkv_bufs", None), self.target_pos_emb.table(),
                list(self.transformer._bf16._d.values()), score_idx]
        ent = dict(key=key, graph=graph, x=x, ids=ids, logits=logits, score_rows=score_rows,
                   launches=_lib.launch_count() - n0, keep=keep)
        cache[nb] = ent
        return ent

    def _mask_predict_device(self, control_emb, steps, preserve, t_overlap, mp_config, long_mode, dynamic):
        """Device-resident mask-predict (the 'batched' sampling mode): every sample and every beam advances in the same
        forward, both random draws of an iteration are one kernel each (csrc/sampling.cu), beam scoring, beam choice and the
        dynamic-stop bookkeeping stay on the device - no host synchronisation anywhere in the loop.  Same algorithm and
        distributions as the reference loop (dalle_bert.py:618-711); its own Philox streams, seeded from torch's CPU
        generator (so torch.manual_seed makes it reproducible), instead of torch's call order.  A sample whose dynamic stop
        has fired (t - tmax >= 5) is frozen rather than left out: the loop always runs Tmax iterations."""
        dev = control_emb.device
        nb, csl, D = control_emb.shape
        Ttot = self.target_seq_len
        MASK = self.image_token_lut["[MASK]"]
        N, pmask, ptok = self._preserve_setup(nb, preserve, t_overlap, long_mode, dev)
        Tmax = mp_config["T"] if steps <= 0 else steps
        Bm = int(mp_config["B"])
        n, temp = mask_predict_schedules(N, mp_config)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # CPU generator: no device sync
        scoring = dynamic or Bm > 1
        use_graph = self._use_cuda_graph(dev)
        from . import _lib as _lib_mod

        def make_runner(batch, control):
            if use_graph:
                fg = self._forward_graph(batch, dev)
                fg["x"][:, :csl].copy_(control)

                def run(ids_in):
                    fg["ids"].copy_(ids_in)
                    fg["graph"].replay()
                    _lib_mod.add_launch_count(fg["launches"])
                    return fg["logits"], fg["score_rows"]
                return run
            xb = torch.empty(batch, self.total_seq_len, D, device=dev, dtype=torch.float32)
            xb[:, :csl].copy_(control)
            sidx = torch.tensor([self.rel_tok_index, self.vid_tok_index], device=dev)

            def run(ids_in):
                ops.embed_gather(xb, [self._target_segment(ids_in)])
                out = self.transformer(xb)
                return (self._head(out[:, csl:].reshape(batch * Ttot, D), self.to_logits).view(batch, Ttot, -1),
                        out.index_select(1, sidx).contiguous())
            return run

        run1 = make_runner(nb, control_emb)
        tok_in = torch.where(pmask, ptok, torch.full_like(ptok, MASK))
        logits, _ = run1(tok_in)
        Y = torch.empty(nb, Ttot, device=dev, dtype=torch.float32)
        I_tok = torch.empty(nb, Ttot, device=dev, dtype=torch.long)
        ops.mp_sample(logits, Y, I_tok, seed, 0, noise_scale=temp[0])
        I_tok = torch.where(pmask, ptok, I_tok)
        if Tmax <= 1:
            if dynamic:
                raise RuntimeError("mask_predict: dynamic=True needs at least 2 steps (the reference returns None here)")
            return I_tok
        runB = run1 if Bm == 1 else make_runner(nb * Bm, control_emb.repeat_interleave(Bm, 0))
        pm_row = pmask[0].contiguous()
        Imax = I_tok
        if dynamic:
            Smax = torch.zeros(nb, device=dev)
            tmax = torch.zeros(nb, device=dev, dtype=torch.long)
            active = torch.ones(nb, device=dev, dtype=torch.bool)
            ar = torch.arange(nb, device=dev)
        elif Bm > 1:
            ar = torch.arange(nb, device=dev)
        for t in range(1, Tmax):
            k = max(N - n[t - 1], 1)
            keep, ids_in = ops.mp_keep(Y, pm_row, I_tok, k, MASK, seed, 2 * t - 1, beams=Bm)
            logits, score_rows = runB(ids_in)
            if not scoring:
                # tokens that were not kept get a new draw; kept ones stay (dalle_bert.py:682-684)
                ops.mp_sample(logits, Y, I_tok, seed, 2 * t, noise_scale=temp[t], skip=keep)
                Imax = I_tok
                continue
            Ynew = torch.empty(nb * Bm, Ttot, device=dev, dtype=torch.float32)
            Inew = torch.empty(nb * Bm, Ttot, device=dev, dtype=torch.long)
            ops.mp_sample(logits, Ynew, Inew, seed, 2 * t, noise_scale=temp[t], skip=keep)
            keepv, Ynew, Inew = keep.view(nb, Bm, Ttot), Ynew.view(nb, Bm, Ttot), Inew.view(nb, Bm, Ttot)
            YB, IB = [], []
            for j in range(Bm):  # beam j starts from beam j-1's state (the reference's cumulative update, :676-684)
                Y = torch.where(keepv[:, j], Y, Ynew[:, j])
                I_tok = torch.where(keepv[:, j], I_tok, Inew[:, j])
                YB.append(Y)
                IB.append(I_tok)
            s_rel = torch.sigmoid(self._head_scalar(score_rows[:, 0].contiguous(), self.to_logits_rel))
            s_vid = torch.sigmoid(self._head_scalar(score_rows[:, 1].contiguous(), self.to_logits_vid))
            S = (0.5 * s_rel + 0.5 * s_vid).view(nb, Bm)
            jmax = S.argmax(dim=1)
            if Bm > 1:
                Y = torch.stack(YB, 1)[ar, jmax]
                I_tok = torch.stack(IB, 1)[ar, jmax]
            if dynamic:
                Sbest = S.gather(1, jmax.view(nb, 1)).view(nb)
                better = (Sbest > Smax) & active
                Smax = torch.where(better, Sbest, Smax)
                tmax = torch.where(better, torch.full_like(tmax, t), tmax)
                Imax = torch.where(better.view(nb, 1), I_tok, Imax)
                active = active & ((t - tmax) < 5)
            else:
                Imax = I_tok
        return Imax

    def _mask_predict_batched(self, control_emb, steps, preserve, t_overlap, mp_config, long_mode):
        """Throughput variant of mask_predict: every sample advances in the same [B,S,D] forward.  Same algorithm
        per sample (beam 1, static schedule); RNG draws are made for the whole batch at once, so ids differ from
        the sample-serial order under the same seed while following the same distribution."""
        dev = control_emb.device
        nb, csl, D = control_emb.shape
        Ttot = self.target_seq_len
        MASK = self.image_token_lut["[MASK]"]
        N, pmask, ptok = self._preserve_setup(nb, preserve, t_overlap, long_mode, dev)
        Tmax = mp_config["T"] if steps <= 0 else steps
        n, temp = mask_predict_schedules(N, mp_config)
        valid_idx = torch.arange(Ttot, device=dev)[~pmask[0]]
        # All Tmax forwards have the same shapes and read the same weights: the chain embed -> 12 layers -> head (~90
        # launches) is captured ONCE per (batch, weights version) into a CUDA graph and replayed; only the token ids
        # change, through a static buffer.  Removes the launch gaps between ~1800 short dependent kernels per call.
        fg = self._forward_graph(nb, dev) if self._use_cuda_graph(dev) else None
        if fg is not None:
            from . import _lib as _lib_mod
            x = fg["x"]
            x[:, :csl].copy_(control_emb)

            def run(ids_in, t):
                fg["ids"].copy_(ids_in)
                fg["graph"].replay()
                _lib_mod.add_launch_count(fg["launches"])
                return self._sample_multinomial(fg["logits"], temp[t])
        else:
            x = torch.empty(nb, self.total_seq_len, D, device=dev, dtype=torch.float32)
            x[:, :csl].copy_(control_emb)

            def run(ids_in, t):
                ops.embed_gather(x, [self._target_segment(ids_in)])
                out = self.transformer(x)
                logits = self._head(out[:, csl:].reshape(nb * Ttot, D), self.to_logits).view(nb, Ttot, -1)
                return self._sample_multinomial(logits, temp[t])

        tok_in = torch.where(pmask, ptok, torch.full_like(ptok, MASK))
        Y, I_new = run(tok_in, 0)
        I_tok = torch.where(pmask, ptok, I_new)
        for t in range(1, Tmax):
            k = max(N - n[t - 1], 1)
            kept = torch.multinomial(Y[:, valid_idx], k, replacement=False)  # [nb, k] indices into valid_idx
            keep = torch.zeros(nb, Ttot, dtype=torch.bool, device=dev)
            keep.scatter_(1, valid_idx[kept], True)
            keep |= pmask
            Y_new, I_new = run(torch.where(keep, I_tok, torch.full_like(I_tok, MASK)), t)
            Y = torch.where(keep, Y, Y_new)
            I_tok = torch.where(keep, I_tok, I_new)
        return I_tok
