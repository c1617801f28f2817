"""CLIP scoring of generated frames on the same kernels (the part that needs no TensorFlow):
`clip_similarity(model, tokenizer, image, description)` of `utils/utils.py:62-85` and the ViT `CLIP` model it calls
(`mmvid_pytorch/transformers/clip_model.py:250-296` VisualTransformer, `:298-420` CLIP.encode_image / encode_text).

Same state-dict keys as the reference's `CLIP` (`visual.conv1.weight`, `visual.class_embedding`, `visual.positional_embedding`,
`visual.ln_pre.*`, `visual.transformer.resblocks.N.*`, `visual.ln_post.*`, `visual.proj`, `transformer.resblocks.N.*`,
`token_embedding.weight`, `positional_embedding`, `ln_final.*`, `text_projection`, `logit_scale`), so `build_model`'s state dict
(`clip_model.py:461-515`) or the TorchScript archive's loads directly.  Only the ViT variants are built (MMVID ships and scores
with ViT-B/32); heads = width // 64 as in `build_model`.

Every FLOP runs in libmmvid_b200: the patch embedding is a GEMM over unfolded 32 x 32 patches, the two stacks are
`OpenAICLIPTransformer` (visual: no mask, text: causal), LayerNorm / projections are the library's kernels.  Frames stay on the
GPU: `clip_similarity` takes the `[N, 3, H, W]` tensor `generate_images` / `vae.decode` returned.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import FP32, H16, PRECISIONS, TF32
from .transformer import OpenAICLIPTransformer, Transformer


class VisualTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution, self.patch_size, self.output_dim = input_resolution, patch_size, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)  # parameter container
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))


class CLIP(nn.Module):
    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length, vocab_size,
                 transformer_width, transformer_heads, transformer_layers, precision="tf32"):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError("ModifiedResNet CLIP variants (MMVID scores with ViT-B/32)")
        assert vision_width % 64 == 0 and transformer_width % 64 == 0
        assert transformer_heads == transformer_width // 64, "build_model: heads = width // 64 (clip_model.py:488)"
        self.context_length_, self.vocab_size = context_length, vocab_size
        self.visual = VisualTransformer(image_resolution, vision_patch_size, vision_width, vision_layers, vision_width // 64,
                                        embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width).normal_(std=0.01))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim).normal_(std=transformer_width ** -0.5))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        # the TorchScript archive exposes these two as tensors (utils/utils.py:63-64 calls .item() on them)
        self.register_buffer("input_resolution", torch.tensor(image_resolution), persistent=False)
        self.register_buffer("context_length", torch.tensor(context_length), persistent=False)
        self.precision = precision
        # kernel runners around the two parameter stacks (kept out of the module tree: the parameters stay under the
        # reference's key names)
        n_vis = (image_resolution // vision_patch_size) ** 2 + 1
        vis = OpenAICLIPTransformer(n_vis, "openai_clip_visual", model_path=None, causal=False, width=vision_width,
                                    layers=vision_layers, precision=precision)
        txt = OpenAICLIPTransformer(context_length, "openai_clip_text", model_path=None, causal=True, mask_type="causal",
                                    width=transformer_width, layers=transformer_layers, precision=precision)
        vis.transformer, txt.transformer = self.visual.transformer, self.transformer
        self._runners = (vis, txt)

    def _prec(self):
        return PRECISIONS[self.precision] if isinstance(self.precision, str) else self.precision

    def _lin_prec(self):
        return FP32 if self._prec() == FP32 else TF32  # patch embedding / projections: fp32 operands, tf32 at most

    @torch.no_grad()
    def encode_image(self, image):
        """image: [N, 3, R, R], already normalised (utils/utils.py:68-71) -> [N, embed_dim] (clip_model.py:275-296)."""
        v = self.visual
        N, _, R, _ = image.shape
        p, g = v.patch_size, R // v.patch_size
        W = v.conv1.weight.shape[0]
        # stride-p conv with a p x p kernel == GEMM over the unfolded patches, (c, ky, kx) fastest like conv1.weight.view(W, -1)
        patches = image.float().reshape(N, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(N * g * g, 3 * p * p).contiguous()
        emb = ops.linear(patches, v.conv1.weight.detach().view(W, -1), None, precision=self._lin_prec()).view(N, g * g, W)
        x = torch.cat([v.class_embedding.detach().expand(N, 1, W), emb], dim=1) + v.positional_embedding.detach()
        x = ops.layernorm(x.contiguous(), v.ln_pre.weight, v.ln_pre.bias, 1e-5)
        self._runners[0].precision = self.precision
        x = self._runners[0](x)
        x = ops.layernorm(x[:, 0].contiguous(), v.ln_post.weight, v.ln_post.bias, 1e-5)
        return ops.linear(x, v.proj.detach().t().contiguous(), None, precision=self._lin_prec())

    @torch.no_grad()
    def encode_text(self, text):
        """text: long [N, context_length] -> [N, embed_dim]; features of the EOT position = argmax id (clip_model.py:399-414)."""
        x = (self.token_embedding.weight.detach()[text] + self.positional_embedding.detach()).contiguous()
        self._runners[1].precision = self.precision
        x = self._runners[1](x)
        rows = x[torch.arange(x.shape[0], device=x.device), text.argmax(dim=-1)].contiguous()
        rows = ops.layernorm(rows, self.ln_final.weight, self.ln_final.bias, 1e-5)  # LayerNorm is per row: select first
        return ops.linear(rows, self.text_projection.detach().t().contiguous(), None, precision=self._lin_prec())

    def forward(self, image, text, **kwargs):
        """clip_model.py:416-431: (logits_per_image, logits_per_text)."""
        fi, ft = self.encode_image(image), self.encode_text(text)
        fi, ft = fi / fi.norm(dim=-1, keepdim=True), ft / ft.norm(dim=-1, keepdim=True)
        logits = self.logit_scale.detach().exp() * fi @ ft.t()
        return logits, logits.t()


def clip_similarity(model, tokenizer, image, description):
    """utils/utils.py:62-85, frames and features staying on the GPU until the final [N] similarities."""
    res, ctx = int(model.input_resolution.item()), int(model.context_length.item())
    if image.shape[2] != res:
        image = F.interpolate(image, (res, res))
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=image.device)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=image.device)
    image_input = (image - mean[:, None, None]) / std[:, None, None]
    text_input = tokenizer.tokenize(description, ctx, truncate_text=True).to(image.device)
    fi, ft = model.encode_image(image_input).float(), model.encode_text(text_input).float()
    fi, ft = fi / fi.norm(dim=-1, keepdim=True), ft / ft.norm(dim=-1, keepdim=True)
    return (ft * fi).sum(1).cpu().numpy()
