"""Builders shared by the tests: the CUDA modules loaded with the SAME synthetic weights the golden generator
gave the reference (mmvid_b200.synth is key-based, so equal state-dict keys => equal tensors)."""
import torch

from cases import codebook_std, fixture_path
from mmvid_b200 import synth


def load_fixture(name):
    return torch.load(fixture_path(name), map_location="cpu", weights_only=False)


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def vae_state_dict(vae_module, seed):
    return synth.fill_state_dict(vae_module, seed, codebook_std=codebook_std())


def build_vae(image_size, seed, device="cuda", precision="fp32"):
    from mmvid_b200.vae import VQGanVAE1024
    vae = VQGanVAE1024(vae_path=None, image_size=image_size, precision=precision)
    vae.image_size = image_size
    sd = vae_state_dict(vae, seed)
    vae.load_state_dict(sd, strict=True)
    return vae.to(device).eval(), sd


def full_state_dict(model, cfg, vae_sd, cvae_sd):
    sd = synth.fill_state_dict(model, cfg["seed"], codebook_std=codebook_std())
    for k, v in vae_sd.items():
        sd["vae." + k] = v.clone()
    if cvae_sd is not None:
        for k, v in cvae_sd.items():
            sd["cvae." + k] = v.clone()
    return sd


def build_bert(cfg, device="cuda", precision="fp32", sampling_mode="reference"):
    from mmvid_b200.dalle_bert import BERT
    vae, vae_sd = build_vae(cfg["image_size"], cfg["seed"] + 1000, device="cpu")
    cvae, cvae_sd = (build_vae(cfg["image_size"], cfg["seed"] + 2000, device="cpu") if cfg["cvae"] else (None, None))
    model = BERT(dim=cfg["dim"], vae=vae, cvae=cvae, num_text_tokens=cfg["vocab"], text_seq_len=cfg["text_seq_len"],
                 which_transformer="openai_clip_visual", num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"],
                 openai_clip_path=None, transformer_layers=cfg["layers"], precision=precision,
                 sampling_mode=sampling_mode)
    sd = full_state_dict(model, cfg, vae_sd, cvae_sd)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), sd


def build_artv(cfg, device="cuda", precision="fp32", sampling_mode="reference"):
    from mmvid_b200.dalle_artv import DALLE
    vae, vae_sd = build_vae(cfg["image_size"], cfg["seed"] + 1000, device="cpu")
    cvae, cvae_sd = build_vae(cfg["image_size"], cfg["seed"] + 2000, device="cpu")
    model = DALLE(dim=cfg["dim"], vae=vae, cvae=cvae, num_text_tokens=cfg["vocab"], text_seq_len=cfg["text_seq_len"],
                  which_transformer="openai_clip_visual", num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"],
                  openai_clip_path=None, transformer_layers=cfg["layers"], precision=precision,
                  sampling_mode=sampling_mode)
    sd = full_state_dict(model, cfg, vae_sd, cvae_sd)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), sd


def bert_spec(cfg):
    from oracle import mmvid_oracle as O
    return O.BertSpec(dim=cfg["dim"], text_seq_len=cfg["text_seq_len"], num_text_tokens=cfg["vocab"],
                      num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"], image_size=cfg["image_size"],
                      has_cvae=cfg["cvae"])


def artv_spec(cfg):
    from oracle import mmvid_oracle as O
    return O.ArtvSpec(dim=cfg["dim"], text_seq_len=cfg["text_seq_len"], num_text_tokens=cfg["vocab"],
                      num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"], image_size=cfg["image_size"])


def to_device(sd, device):
    return {k: v.to(device) for k, v in sd.items()}
