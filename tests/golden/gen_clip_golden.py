"""Golden vectors for mmvid_b200.clip_score: the UNMODIFIED reference CLIP (/root/reference/mmvid_pytorch/transformers/
clip_model.py, class CLIP with a ViT tower) run on CPU in fp32 on seeded weights and inputs - image features, text features
and the `clip_similarity` arithmetic of utils/utils.py:62-85.  Run in the build container only:

    python tests/golden/gen_clip_golden.py        # writes tests/golden/clip_small.pt
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(embed_dim=64, image_resolution=96, vision_layers=2, vision_width=64, vision_patch_size=32, context_length=16,
           vocab_size=300, transformer_width=64, transformer_heads=1, transformer_layers=2)

if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("ref_clip_model", "/root/reference/mmvid_pytorch/transformers/clip_model.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(11)
    model = ref.CLIP(**CFG).float().eval()
    # LayerNorm affine / biases off their identity initialisation so that every parameter matters
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.ndim == 1 and "class_embedding" not in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    image = torch.rand(3, 3, 96, 96, generator=g)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073])
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711])
    image_input = (image - mean[:, None, None]) / std[:, None, None]
    text = torch.zeros(3, 16, dtype=torch.long)
    for i, n_tok in enumerate((5, 16, 9)):
        text[i, :n_tok] = torch.randint(1, 290, (n_tok,), generator=g)
        text[i, n_tok - 1] = 299  # EOT = highest id: encode_text picks argmax
    with torch.no_grad():
        fi = model.encode_image(image_input).float()
        ft = model.encode_text(text).float()
        li, lt = model(image_input, text)
    sim = ((ft / ft.norm(dim=-1, keepdim=True)) * (fi / fi.norm(dim=-1, keepdim=True))).sum(1)
    out = {"cfg": CFG, "state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "image": image, "text": text,
           "image_features": fi, "text_features": ft, "similarity": sim, "logits_per_image": li}
    torch.save(out, os.path.join(HERE, "clip_small.pt"))
    print("wrote clip_small.pt", {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)}, "similarity", sim.tolist())
