"""B200-native mirror of `mmvid_pytorch/dalle_artv.py:103-542` (class DALLE, the ART-V autoregressive baseline).

Same constructor keywords / attributes / state-dict keys (`text_emb`, `text_pos_emb` [L+1], `image_emb` [1024],
`image_pos_emb.weights_*`, `visual_emb` [1024+V*n], `visual_pos_emb.module_list.*`, `special_emb`,
`estimation_pos_emb`, `transformer.transformer.resblocks.*`, `to_logits.{0,1}`), same `forward` (masked logits) and
`generate_images` (returns `(images, [], None)`) contracts.

What is new: `generate_images` keeps a per-layer K/V cache.  The reference re-runs the whole prefix through all
layers, re-encodes the visual frames with the VQGAN and evaluates a 51 k-wide vocabulary head at every
position for every sampled token (dalle_artv.py:258-281, :464, :505); with a causal mask and additive absolute
position embeddings the cache is exact, so one step is: embed 1 token -> 12 x (LN, QKV GEMV, cache append,
single-query attention, out-proj, LN, MLP) -> LN + 1024-column image head.  Sampling (`top_k` -> softmax ->
`torch.multinomial` over the full `total_tokens` axis) keeps the reference's RNG shapes in 'reference' mode.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import FP32, PRECISIONS
from .dalle_bert import eval_decorator, exists, set_requires_grad
from .modules import AxialPositionalEmbedding, AxialPositionalEmbeddingList
from .transformer import OpenAICLIPTransformer


def is_empty(t):
    return t.nelement() == 0


def top_k(logits, thres=0.5):
    """dalle_artv.py:61-67."""
    k = max(int((1 - thres) * logits.shape[-1]), 1)
    val, ind = torch.topk(logits, k)
    probs = torch.full_like(logits, float("-inf"))
    probs.scatter_(1, ind, val)
    return probs


class DALLE(nn.Module):
    def __init__(self, *, dim, vae, cvae=None, num_text_tokens=10000, text_seq_len=256, loss_img_weight=7, stable=False,
                 which_transformer="none", num_visuals=1, num_targets=1, **kwargs):
        super().__init__()
        assert num_visuals > 0
        if stable:
            raise NotImplementedError
        num_image_tokens = vae.num_tokens
        fmap = vae.image_size // (2 ** vae.num_layers)
        n = fmap ** 2
        self.insert_sep = False
        num_text_tokens = num_text_tokens + text_seq_len           # unique pad ids (dalle_artv.py:132)
        num_visual_tokens = num_image_tokens + n * num_visuals     # :133
        self.text_emb = nn.Embedding(num_text_tokens, dim)
        self.image_emb = nn.Embedding(num_image_tokens, dim)
        self.text_pos_emb = nn.Embedding(text_seq_len + 1, dim)    # +1 for <bos>
        shape = (fmap, fmap) if num_targets == 1 else (num_targets, fmap, fmap)
        self.image_pos_emb = AxialPositionalEmbedding(dim, axial_shape=shape)
        self.visual_emb = nn.Embedding(num_visual_tokens, dim)
        self.visual_pos_emb = AxialPositionalEmbeddingList(dim, num_visuals, axial_shape=(fmap, fmap))

        self.dim = dim
        self.num_text_tokens = num_text_tokens
        self.num_image_tokens = num_image_tokens
        self.num_visual_tokens = num_visual_tokens
        self.num_control_tokens = num_text_tokens + num_visual_tokens
        self.text_seq_len = text_seq_len
        self.image_seq_len = n
        self.target_seq_len = n * num_targets
        self.visual_seq_len = n * num_visuals
        self.control_seq_len = text_seq_len + self.visual_seq_len
        self.num_visuals, self.num_targets = num_visuals, num_targets
        self.image_fmap_size = fmap
        self.special_token_lut = {"[REL]": 0, "[ST1]": 1, "[ST2]": 2, "[ST3]": 3}
        self.num_special_tokens = 4
        self.num_estimation_tokens = 2
        self.special_emb = nn.Embedding(4, dim)              # allocated by the reference, unused in forward
        self.estimation_pos_emb = nn.Embedding(2, dim)
        self.total_tokens = num_text_tokens + num_image_tokens + num_visual_tokens
        self.total_seq_len = text_seq_len + self.target_seq_len + self.visual_seq_len
        self.vae, self.cvae = vae, cvae
        set_requires_grad(self.vae, False)
        set_requires_grad(self.cvae, False)
        self.which_transformer = which_transformer
        if not which_transformer.startswith("openai_clip"):
            raise NotImplementedError
        self.transformer = OpenAICLIPTransformer(self.total_seq_len, which_transformer,
                                                 model_path=kwargs.get("openai_clip_path"), width=dim,
                                                 layers=kwargs.get("transformer_layers"),
                                                 precision=kwargs.get("precision", "tf32"))
        self.stable = False
        self.to_logits = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, self.total_tokens))
        self.loss_vis_weight = 1.0
        self.loss_img_weight = loss_img_weight
        self.sampling_mode = kwargs.get("sampling_mode", "reference")

    @property
    def precision(self):
        return self.transformer.precision

    @precision.setter
    def precision(self, p):
        assert p in PRECISIONS
        self.transformer.precision = p

    # logits_mask (dalle_artv.py:215-227, persistent=False): row r may emit only its own vocabulary block.
    def _allowed_range(self, row):
        if row < self.text_seq_len:
            return 0, self.num_text_tokens
        if row < self.control_seq_len:
            return self.num_text_tokens, self.num_control_tokens
        return self.num_control_tokens, self.total_tokens

    # ------------------------------------------------------------------------------------------ token helpers
    def get_image_tokens(self, image, reshape=True, insert_sep=False, which_vae="vae"):
        vae = self.cvae if (which_vae == "cvae" and self.cvae is not None) else self.vae
        if isinstance(image, list):
            image = torch.stack(image, dim=1)
        if image.dim() == 4:
            image = image.unsqueeze(1)
        if image.dim() == 5:
            b, t, c, h, w = image.shape
            assert (c, h, w) == (3, vae.image_size, vae.image_size), \
                f"invalid image of dimensions {image.shape} passed in during training"
            ids = vae.get_codebook_indices(image.reshape(b * t, c, h, w))
            image = ids.view(b, t * ids.shape[1]) if reshape else ids
        return image

    @torch.no_grad()
    def recon_images(self, images, which_vae="vae"):
        vae = self.cvae if (which_vae == "cvae" and self.cvae is not None) else self.vae
        return vae.decode(self.get_image_tokens(images, reshape=False, which_vae=which_vae))

    def random_erase_codebook(self, image, eraser, erase_half=False):
        f = self.image_fmap_size
        grid = image.view(image.shape[0], -1, f, f)
        if erase_half:
            grid[:, :, f // 2:, :] = -1
            out = grid
        else:
            out = torch.stack([eraser(c) for c in grid], dim=0)
        return out.reshape(image.shape[0], -1)

    def erase_codebook_face(self, image, vc_mode, face_mode=None):
        """dalle_artv.py:373-416 (erased positions carry -1 -> unique pad ids)."""
        import random
        import numpy as np
        f = self.image_fmap_size
        grid = image.view(image.shape[0], -1, f, f)

        def keep_only(windows):
            out = torch.full_like(grid, -1)
            for tsel, rs, cs in windows:
                out[:, tsel, rs, cs] = grid[:, tsel, rs, cs]
            return out

        every = slice(None)
        if vc_mode == "face_8x8":
            if face_mode is None:
                face_mode = "eyes_nose" if random.random() < 0.5 else "mouth"
            win = (slice(2, 5), slice(1, 7)) if face_mode == "eyes_nose" else (slice(5, 7), slice(2, 6))
            grid = keep_only([(every, *win)])
        elif vc_mode == "face2_8x8":
            grid = keep_only([(slice(0, 1), every, every), (every, slice(2, 6), slice(2, 6))])
        elif vc_mode in ("mask_8x8", "mask2_8x8"):
            strategy = int(np.random.choice([1, 2, 3], p=[0.5, 0.25, 0.25])) if face_mode is None else 3
            if strategy == 3:
                grid = keep_only([(every, slice(1, 7), slice(1, 7))])
            # strategy 2 of the reference computes a masked copy but never assigns it (dalle_artv.py:401-403)
        elif vc_mode == "shape_4x4":
            grid[:, :, 1:3, 1:3] = -1
        else:
            raise NotImplementedError
        return grid.reshape(image.shape[0], -1)

    def _visual_ids(self, visual, B, dev, erase_visual, erase_visual_half, vc_mode, face_mode):
        if exists(visual) and not is_empty(visual):
            ids = self.get_image_tokens(visual, insert_sep=False, which_vae="cvae")
            if erase_visual:
                import torchvision.transforms as T
                eraser = T.RandomErasing(p=1, scale=(0.4, 0.8), ratio=(0.5, 2), value=-1)
                ids = self.random_erase_codebook(ids, eraser, erase_visual_half)
            if vc_mode is not None:
                ids = self.erase_codebook_face(ids, vc_mode, face_mode)
            return ids
        return torch.full((B, self.visual_seq_len), -1, dtype=torch.long, device=dev)

    # ------------------------------------------------------------------------------------------ embeddings
    def _prefix_segments(self, text, visual_ids):
        bos_text = F.pad(torch.where(text == 0, torch.arange(self.text_seq_len, device=text.device) +
                                     (self.num_text_tokens - self.text_seq_len), text), (1, 0), value=0)
        segs = [dict(ids=bos_text.contiguous(), seq_off=0, table=self.text_emb.weight.detach(),
                     pos=self.text_pos_emb.weight.detach()),
                dict(ids=visual_ids.contiguous(), seq_off=self.text_seq_len + 1, table=self.visual_emb.weight.detach(),
                     pos=self.visual_pos_emb.table(),
                     pad=(-1, self.num_visual_tokens - self.visual_seq_len))]
        return segs

    def _head_rows(self, rows, col_lo, col_hi):
        """to_logits (LN + Linear) restricted to vocabulary columns [col_lo, col_hi)."""
        prec = PRECISIONS[self.precision]
        ln, lin = self.to_logits[0], self.to_logits[1]
        if rows.shape[0] <= 16:
            h = ops.layernorm(rows, ln.weight, ln.bias, 1e-5)
            return ops.linear_small_m(h, lin.weight.detach()[col_lo:col_hi], lin.bias.detach()[col_lo:col_hi])
        h = ops.layernorm(rows, ln.weight, ln.bias, 1e-5, out_dtype=ops.act_dtype(prec))
        w = self.transformer._w(lin.weight, prec)[col_lo:col_hi]
        return ops.linear(h, w, lin.bias.detach()[col_lo:col_hi], precision=prec)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, text, visual=None, target=None, return_loss=False, erase_visual=False, erase_visual_half=False,
                vc_mode=None, face_mode=None, visual_aug_mode=None, **kwargs):
        """Masked logits [B, seq, total_tokens] (dalle_artv.py:418-515); `target` = image ids [B, t] or frames."""
        assert text.shape[-1] == self.text_seq_len, \
            f"the length {text.shape[-1]} of the text tokens you passed in does not have the correct length ({self.text_seq_len})"
        if return_loss:
            raise NotImplementedError("training losses / backward are scheduled after the inference path")
        if visual_aug_mode is not None and exists(visual) and not is_empty(visual):
            from .augment import augment_visual
            visual = augment_visual(visual, visual_aug_mode)  # dalle_artv.py:460-463
        B, dev = text.shape[0], text.device
        with torch.no_grad():
            visual_ids = self._visual_ids(visual, B, dev, erase_visual, erase_visual_half, vc_mode, face_mode)
            segs = self._prefix_segments(text, visual_ids)
            seq = self.text_seq_len + 1 + self.visual_seq_len
            image = None
            if exists(target) and not is_empty(target):
                image = self.get_image_tokens(target)
                segs.append(dict(ids=image.contiguous(), seq_off=seq, table=self.image_emb.weight.detach(),
                                 pos=self.image_pos_emb.table()))
                seq += image.shape[1]
            x = torch.empty(B, seq, self.dim, device=dev, dtype=torch.float32)
            ops.embed_gather(x, segs)
            if seq > self.total_seq_len:  # :496-498
                seq -= 1
                x = x[:, :seq].contiguous()
            out = self.transformer(x)
            logits = self._head_rows(out.reshape(B * seq, self.dim), 0, self.total_tokens).view(B, seq, -1)
            neg = -torch.finfo(logits.dtype).max
            for lo_row, hi_row in ((0, self.text_seq_len), (self.text_seq_len, self.control_seq_len),
                                   (self.control_seq_len, seq)):
                if lo_row >= seq:
                    break
                hi_row = min(hi_row, seq)
                a, b = self._allowed_range(lo_row)
                logits[:, lo_row:hi_row, :a] = neg
                logits[:, lo_row:hi_row, b:] = neg
        return logits

    # ------------------------------------------------------------------------------------------ sampling
    def generate_tokens(self, text, visual=None, max_new=None, **kwargs):
        """The sampling loop of generate_images alone: long [B, max_new] image-token ids (no VQGAN decode)."""
        return self.generate_images(text, visual=visual, max_new=max_new if max_new is not None else self.target_seq_len,
                                    **kwargs)[2]

    @torch.no_grad()
    @eval_decorator
    def generate_images(self, text, *, clip=None, visual=None, mask=None, filter_thres=0.5, temperature=1.0,
                        erase_visual=False, vc_mode=None, face_mode=None, return_tokens=False, **kwargs):
        """dalle_artv.py:236-304 with a KV cache.  Returns (images, [], None) like the reference."""
        text = text[:, : self.text_seq_len]
        B, dev, D = text.shape[0], text.device, self.dim
        H = self.transformer.transformer.heads
        blocks = list(self.transformer.transformer.resblocks)
        visual_ids = self._visual_ids(visual, B, dev, erase_visual, True, vc_mode, face_mode)
        P = self.text_seq_len + 1 + self.visual_seq_len
        S_max = P + self.target_seq_len
        x = torch.empty(B, P, D, device=dev, dtype=torch.float32)
        ops.embed_gather(x, self._prefix_segments(text, visual_ids))
        # ---- decode implementation.  'stream' (default in the tensor-core precisions, B <= 8): ONE persistent launch per
        # token on 16-bit weights and a 16-bit K/V cache (decode_stream.cu).  fp32 precision keeps the fp32 kernels that
        # reproduce the reference's sampled ids bit for bit: 'fused' (5 PDL-chained launches / layer, decode_pdl.cu),
        # 'persistent' (cooperative, fp32), 'native' (8 launches / layer).
        prec = PRECISIONS[self.precision]
        impl = os.environ.get("MMVID_ARTV_DECODE", getattr(self, "decode_impl", None) or ("fused" if prec == 0 else "stream"))
        stream = impl == "stream" and B <= 8 and prec != 0
        if impl == "stream" and not stream:
            impl = "fused"
        h16 = torch.bfloat16 if prec == 2 else torch.float16  # tf32 mode: fp16 copies carry the same 10-bit mantissa
        cache_dt = h16 if stream else torch.float32
        # ---- prefill: full causal forward over the prefix, K/V of every layer captured into the caches
        kc = [torch.zeros(B, H, S_max, 64, device=dev, dtype=cache_dt) for _ in blocks]
        vc = [torch.zeros(B, H, S_max, 64, device=dev, dtype=cache_dt) for _ in blocks]
        hid = self.transformer(x, kv_out=(kc, vc))
        lo = self.num_control_tokens
        logits = self._head_rows(hid[:, -1].contiguous(), lo, lo + self.num_image_tokens)
        pos_table = self.image_pos_emb.table()
        out_tokens = torch.empty(B, self.target_seq_len, dtype=torch.long, device=dev)
        k_keep = max(int((1 - filter_thres) * self.total_tokens), 1)
        if k_keep < self.num_image_tokens:
            raise NotImplementedError("filter_thres that prunes inside the image vocabulary")
        xt = torch.empty(B, 1, D, device=dev, dtype=torch.float32)
        native = B <= 16
        ln, lin = self.to_logits[0], self.to_logits[1]
        if stream:
            from . import _lib as L
            lib = L.load()
            w16 = self.transformer._bf16
            layers16 = (L.DecodeLayer16 * len(blocks))()
            keep16 = []
            for li, blk in enumerate(blocks):
                e = layers16[li]
                e.ln1_w, e.ln1_b = blk.ln_1.weight.data_ptr(), blk.ln_1.bias.data_ptr()
                e.ln2_w, e.ln2_b = blk.ln_2.weight.data_ptr(), blk.ln_2.bias.data_ptr()
                e.in_b, e.out_b = blk.attn.in_proj_bias.data_ptr(), blk.attn.out_proj.bias.data_ptr()
                e.fc_b, e.proj_b = blk.mlp.c_fc.bias.data_ptr(), blk.mlp.c_proj.bias.data_ptr()
                ws_ = [w16.get(q, h16) for q in (blk.attn.in_proj_weight, blk.attn.out_proj.weight, blk.mlp.c_fc.weight,
                                                 blk.mlp.c_proj.weight)]
                keep16.append(ws_)
                e.in_w, e.out_w, e.fc_w, e.proj_w = (t.data_ptr() for t in ws_)
                e.kcache, e.vcache = kc[li].data_ptr(), vc[li].data_ptr()
            head_w16 = w16.get(lin.weight, h16)[lo:lo + self.num_image_tokens]
            head_b = lin.bias.detach()[lo:lo + self.num_image_tokens]
            ws16 = torch.zeros(int(lib.mmvid_artv_decode_stream_workspace_floats(B, D, H)), device=dev, dtype=torch.float32)
            logits_buf = torch.empty(B, self.num_image_tokens, device=dev, dtype=torch.float32)
        if native and not stream:
            from . import _lib as L
            import ctypes as C
            lib = L.load()
            layers = (L.DecodeLayer * len(blocks))()
            for li, blk in enumerate(blocks):
                e = layers[li]
                e.ln1_w, e.ln1_b = blk.ln_1.weight.data_ptr(), blk.ln_1.bias.data_ptr()
                e.in_w, e.in_b = blk.attn.in_proj_weight.data_ptr(), blk.attn.in_proj_bias.data_ptr()
                e.out_w, e.out_b = blk.attn.out_proj.weight.data_ptr(), blk.attn.out_proj.bias.data_ptr()
                e.ln2_w, e.ln2_b = blk.ln_2.weight.data_ptr(), blk.ln_2.bias.data_ptr()
                e.fc_w, e.fc_b = blk.mlp.c_fc.weight.data_ptr(), blk.mlp.c_fc.bias.data_ptr()
                e.proj_w, e.proj_b = blk.mlp.c_proj.weight.data_ptr(), blk.mlp.c_proj.bias.data_ptr()
                e.kcache, e.vcache = kc[li].data_ptr(), vc[li].data_ptr()
            ws = torch.zeros(int(lib.mmvid_artv_decode_workspace_floats(B, D, H)), device=dev, dtype=torch.float32)
            persistent = impl == "persistent" and B <= 8 and len(blocks) <= 24
            fused = impl == "fused" and B <= 8
            head_w = lin.weight.detach()[lo:lo + self.num_image_tokens]
            head_b = lin.bias.detach()[lo:lo + self.num_image_tokens]
            logits_buf = torch.empty(B, self.num_image_tokens, device=dev, dtype=torch.float32)
        trace = kwargs.get("logits_trace")  # optional list: receives the [B, 1024] image logits of every step (tests)
        max_new = kwargs.get("max_new")     # stop after this many sampled tokens and return them (parity tests)
        n_steps = self.target_seq_len if max_new is None else min(int(max_new), self.target_seq_len)

        def stream_step(h_buf, pos, pos_dev=None):
            rc = lib.mmvid_artv_decode_stream(layers16, len(blocks), ops._ptr(h_buf), ops._ptr(ws16), ops._ptr(ln.weight),
                                              ops._ptr(ln.bias), ops._ptr(head_w16), ops._ptr(head_b), ops._ptr(logits_buf),
                                              self.num_image_tokens, B, D, H, S_max, pos, ops._ptr(pos_dev),
                                              int(h16 == torch.float16), ops._stream())
            if rc == 1:
                raise RuntimeError("mmvid_artv_decode_stream: shape does not fit the kernel's shared-memory plan; "
                                   "use precision='fp32' (decode_impl='fused')")
            L.check(rc, "artv_decode_stream")

        graphed = (stream and self.sampling_mode != "reference" and trace is None and n_steps > 2
                   and os.environ.get("MMVID_CUDA_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing())
        if graphed:
            # The token loop is a chain of ~8 short launches per token whose host side (Python, ctypes, torch dispatch)
            # costs several times the device time of the one-launch decode step.  One token step - softmax, multinomial,
            # token store, embedding + position gather, decode kernel, step counter - is captured ONCE into a CUDA graph; the
            # step index lives in device memory (t_dev), so every replay is the next token.
            emb_w = self.image_emb.weight.detach()
            t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            logits_buf.copy_(logits)
            saved_logits = logits_buf.clone()
            h_buf = xt.view(B, D)
            inv_t = 1.0 / float(temperature)

            seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # CPU generator (follows torch.manual_seed), no device sync
            y_buf = torch.empty(B, device=dev, dtype=torch.float32)
            tok_buf = torch.empty(B, device=dev, dtype=torch.long)

            def token_step():
                lg = logits_buf * inv_t if temperature != 1.0 else logits_buf
                # fused softmax + categorical draw (csrc/sampling.cu); the Philox offset follows the device step counter
                ops.mp_sample(lg, y_buf, tok_buf, seed, 0, step_dev=t_dev)
                t_idx = t_dev.long()
                out_tokens.scatter_(1, t_idx.view(1, 1).expand(B, 1), tok_buf.view(B, 1))
                h_buf.copy_(emb_w[tok_buf] + pos_table.index_select(0, t_idx))
                stream_step(h_buf, P, t_dev)
                t_dev.add_(1)

            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up: one-time kernel attributes and allocator state; then restore the inputs
                token_step()
                logits_buf.copy_(saved_logits)
                t_dev.zero_()
            cur.wait_stream(side)
            # Manual capture instead of `with torch.cuda.graph(...)`: the context manager empties the caching allocator on
            # entry (cudaFree of every cached block, then cudaMalloc again for the VQGAN decode that follows) - 0.1-0.7 s of
            # host time per generate_images call on a busy box, inside the benchmark's timed region.
            def capture(n_tokens, pool=None):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    g_.capture_begin(*(() if pool is None else (pool,)))
                    try:
                        for _ in range(n_tokens):
                            token_step()
                    finally:
                        g_.capture_end()
                return g_
            side.wait_stream(cur)
            n0 = L.launch_count()
            graph = capture(1)
            per_replay = L.launch_count() - n0
            # ... and GROUP token steps in a second graph: 2047 replays of an 8-node graph are ~2047 host round trips
            # (20-400 us each, depending on how busy the host is - one bench leg measured 1.67 s instead of 0.9 s on a noisy
            # box); with the step counter on the device a graph of GROUP consecutive steps is just as valid.
            GROUP = 16
            todo = n_steps - 1
            graph_n = None
            if todo >= 2 * GROUP:
                graph_n = capture(GROUP, graph.pool())
            cur.wait_stream(side)
            captured = L.launch_count()
            logits_buf.copy_(saved_logits)
            t_dev.zero_()
            while graph_n is not None and todo >= GROUP:
                graph_n.replay()
                todo -= GROUP
            for _ in range(todo):
                graph.replay()
            L.add_launch_count(per_replay * (n_steps - 1) - (captured - n0))  # captures were counted as launches
            # the last token only needs the sampling half of the step
            lg = logits_buf * inv_t if temperature != 1.0 else logits_buf
            ops.mp_sample(lg, y_buf, tok_buf, seed, 0, step_dev=t_dev)
            out_tokens[:, n_steps - 1] = tok_buf
            n_steps_loop = 0
        else:
            n_steps_loop = n_steps
        for t in range(n_steps_loop):
            if trace is not None:
                trace.append(logits.clone())
            # top_k keeps >= 1024 entries (k = 25888 at the default 0.5), so only the masked logits (-FLT_MAX -> prob 0)
            # are affected: softmax over the 1024 image logits is the whole distribution (dalle_artv.py:274-276)
            probs_img = ops.softmax_logits(logits / temperature if temperature != 1.0 else logits)
            if self.sampling_mode == "reference":
                full = torch.zeros(B, self.total_tokens, device=dev, dtype=torch.float32)
                full[:, lo:] = probs_img
                sample = torch.multinomial(full, 1) - lo
            else:
                sample = torch.multinomial(probs_img, 1)
            out_tokens[:, t] = sample[:, 0]
            if t == n_steps - 1:
                break
            # ---- one decode step at sequence position P + t
            ops.embed_gather(xt, [dict(ids=sample, seq_off=0, table=self.image_emb.weight.detach(),
                                       pos=pos_table[t:t + 1])])
            h = xt.view(B, D)
            pos = P + t
            if stream:
                stream_step(h, pos)
                logits = logits_buf
                continue
            if native and fused:
                L.check(lib.mmvid_artv_decode_fused(layers, len(blocks), ops._ptr(h), ops._ptr(ws), ops._ptr(ln.weight),
                                                    ops._ptr(ln.bias), ops._ptr(head_w), ops._ptr(head_b),
                                                    ops._ptr(logits_buf), self.num_image_tokens, B, D, H, S_max, pos,
                                                    ops._stream()), "artv_decode_fused")
                logits = logits_buf
                continue
            if native and persistent:
                # ONE cooperative launch: 12 layers + LN + image-logit head, grid barriers between phases
                L.check(lib.mmvid_artv_decode_persistent(layers, len(blocks), ops._ptr(h), ops._ptr(ws), ops._ptr(ln.weight),
                                                         ops._ptr(ln.bias), ops._ptr(head_w), ops._ptr(head_b),
                                                         ops._ptr(logits_buf), self.num_image_tokens, B, D, H, S_max, pos,
                                                         ops._stream()), "artv_decode_persistent")
                logits = logits_buf
                continue
            if native:
                # all 12 layers of this token issued from C (8 launches / layer, no Python in between)
                L.check(lib.mmvid_artv_decode_step(layers, len(blocks), ops._ptr(h), ops._ptr(ws), B, D, H, S_max, pos,
                                                   ops._stream()), "artv_decode_step")
                logits = self._head_rows(h, lo, lo + self.num_image_tokens)
                continue
            for li, blk in enumerate(blocks):
                a = ops.layernorm(h, blk.ln_1.weight, blk.ln_1.bias, 1e-5)
                qkv = ops.linear_small_m(a, blk.attn.in_proj_weight.detach(), blk.attn.in_proj_bias)
                ops.kv_append(qkv, kc[li], vc[li], pos)
                att = ops.decode_attention(qkv, kc[li], vc[li], pos + 1)
                h = ops.linear_small_m(att, blk.attn.out_proj.weight.detach(), blk.attn.out_proj.bias, residual=h)
                a = ops.layernorm(h, blk.ln_2.weight, blk.ln_2.bias, 1e-5)
                a = ops.linear_small_m(a, blk.mlp.c_fc.weight.detach(), blk.mlp.c_fc.bias, act=1)
                h = ops.linear_small_m(a, blk.mlp.c_proj.weight.detach(), blk.mlp.c_proj.bias, residual=h)
            logits = self._head_rows(h, lo, lo + self.num_image_tokens)
        if max_new is not None:
            return None, [], out_tokens[:, :n_steps]
        img_seq = out_tokens.reshape(-1, self.image_seq_len) if self.num_targets > 1 else out_tokens
        images = self.vae.decode(img_seq)
        if self.num_targets > 1:
            images = images.view(B, self.num_targets, *images.shape[1:])
        if return_tokens:
            return images, [], out_tokens
        if exists(clip):
            raise NotImplementedError("CLIP re-ranking is outside the hot path")
        return images, [], None
