"""Generate the golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU fp32.

    python tests/golden/gen_golden.py [case ...]

Runs only in the build container (the reference is not present on the GPU box).  For every case it
 1. builds the reference module (through oracle/ref_shims.py), loads synthetic weights (mmvid_b200.synth),
 2. runs the reference on seeded inputs,
 3. runs oracle/mmvid_oracle.py on the same tensors and ASSERTS agreement (this is what pins the oracle),
 4. writes the reference outputs to tests/golden/<case>.pt (small tensors only).
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import (ARTV_CASES, BERT_CASES, ROOT, TRANSFORMER_CASES, VAE_CASES, codebook_std,  # noqa: E402
                   fixture_path)

sys.path.insert(0, ROOT)
from mmvid_b200 import synth  # noqa: E402
from oracle import mmvid_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402

torch.set_grad_enabled(False)


def relerr(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def build_ref_vae(image_size, seed):
    ref_shims.install()
    from mmvid_pytorch.vae import VQGanVAE1024
    with ref_shims.ref_cwd():
        vae = VQGanVAE1024(vae_path=None, image_size=image_size)
    vae.image_size = image_size  # train.py:182-185
    sd = synth.fill_state_dict(vae, seed, codebook_std=codebook_std())
    vae.load_state_dict(sd, strict=True)
    return vae.eval(), sd


def build_ref_bert(cfg):
    ref_shims.install()
    from mmvid_pytorch.dalle_bert import BERT
    vae, _ = build_ref_vae(cfg["image_size"], cfg["seed"] + 1000)
    cvae = build_ref_vae(cfg["image_size"], cfg["seed"] + 2000)[0] if cfg["cvae"] else None
    clip_sd = synth.clip_checkpoint_state_dict(cfg["dim"], cfg["layers"], seed=cfg["seed"])
    with ref_shims.fake_clip_checkpoint(clip_sd):
        model = BERT(dim=cfg["dim"], vae=vae, cvae=cvae, num_text_tokens=cfg["vocab"],
                     text_seq_len=cfg["text_seq_len"], which_transformer="openai_clip_visual",
                     num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"], openai_clip_path="none")
    sd = synth.fill_state_dict(model, cfg["seed"], codebook_std=codebook_std())
    # vae / cvae weights keep their own seeds (so VAE fixtures and BERT fixtures share them)
    for k, v in vae.state_dict().items():
        sd["vae." + k] = v.clone()
    if cvae is not None:
        for k, v in cvae.state_dict().items():
            sd["cvae." + k] = v.clone()
    model.load_state_dict(sd, strict=True)
    return model.eval(), sd


def bert_spec(cfg):
    return O.BertSpec(dim=cfg["dim"], text_seq_len=cfg["text_seq_len"], num_text_tokens=cfg["vocab"],
                      num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"], image_size=cfg["image_size"],
                      has_cvae=cfg["cvae"])


def gen_transformer(name, cfg):
    ref_shims.install()
    from mmvid_pytorch.transformers.clip_model import OpenAICLIPTransformer
    clip_sd = synth.clip_checkpoint_state_dict(cfg["dim"], cfg["layers"], seed=cfg["seed"])
    with ref_shims.fake_clip_checkpoint(clip_sd):
        m = OpenAICLIPTransformer(cfg["seq"], "openai_clip_visual", model_path="none", causal=True,
                                  mask_type=cfg["mask"], mask_kwargs={"index": list(cfg["index"])})
    sd = synth.fill_state_dict(m, cfg["seed"])
    m.load_state_dict(sd)
    m.eval()
    x = synth.synth_tensor("x", (cfg["batch"], cfg["seq"], cfg["dim"]), cfg["seed"] + 7)
    y_ref = m(x)
    mask = O.build_attention_mask(cfg["seq"], cfg["mask"], cfg["index"])
    y_or = O.transformer_forward(x, sd, "transformer.", mask)
    e = relerr(y_or, y_ref)
    assert e < 2e-6, (name, e)
    torch.save(dict(cfg=cfg, y=y_ref.clone(), oracle_relerr=e), fixture_path(name))
    print(f"{name}: oracle vs reference relerr {e:.2e}")


def gen_vae(name, cfg):
    vae, sd = build_ref_vae(cfg["image_size"], cfg["seed"])
    img = synth.synth_frames(cfg["batch"], 1, cfg["image_size"], cfg["seed"])[:, 0]
    idx_ref = vae.get_codebook_indices(img)
    # pre-quant latent (what VectorQuantizer2 sees) and distance gap statistics
    z_ref = vae.model.quant_conv(vae.model.encoder(2 * img - 1))
    dec_ref = vae.decode(idx_ref)
    rand_codes = synth.synth_codes(cfg["batch"], idx_ref.shape[1], seed=cfg["seed"])
    dec_rand_ref = vae.decode(rand_codes)
    osd = {"model." + k if not k.startswith("model.") else k: v for k, v in sd.items()}
    z_or = O.vae_pre_quant(img, osd)
    idx_or = O.vae_get_codebook_indices(img, osd)
    dec_or = O.vae_decode(idx_ref, osd)
    assert torch.equal(idx_or, idx_ref), name
    ez, ed = relerr(z_or, z_ref), relerr(dec_or, dec_ref)
    assert ez < 1e-5 and ed < 1e-5, (name, ez, ed)
    d = O.vq_distances(z_ref.permute(0, 2, 3, 1).reshape(-1, 256), sd["model.quantize.embedding.weight"])
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    gap = float((top2[:, 1] - top2[:, 0]).min())
    torch.save(dict(cfg=cfg, indices=idx_ref.clone(), z=z_ref.clone(), decoded=dec_ref.clone(),
                    rand_codes=rand_codes, decoded_rand=dec_rand_ref.clone(), min_top2_gap=gap,
                    z_std=float(z_ref.std())), fixture_path(name))
    print(f"{name}: idx exact, z relerr {ez:.2e}, dec relerr {ed:.2e}, min top-2 gap {gap:.3e}, z std {z_ref.std():.3f}")


def gen_bert(name, cfg):
    model, sd = build_ref_bert(cfg)
    spec = bert_spec(cfg)
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"])
    visual = None
    if cfg["num_visuals"] > 0:
        visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5)
    out = dict(cfg=cfg)
    t0 = time.time()
    vc = dict(vc_mode=cfg["vc_mode"], face_mode=cfg.get("face_mode")) if cfg.get("vc_mode") else {}
    control_ref = model(text, visual=visual, return_loss=False, **vc)
    vis_tok = None
    if visual is not None:
        p = "cvae." if cfg["cvae"] else "vae."
        vis_tok = O.vae_get_codebook_indices(visual.reshape(-1, *visual.shape[2:]), O.sub_state_dict(sd, p)).view(B, -1)
        if vc:
            # the reference's own tokens after its own erase hook, to pin the oracle's restatement of it
            ref_tok = model.erase_codebook_face(model.get_image_tokens(visual, which_vae="cvae"), **vc)
            out["visual_tokens_raw"] = vis_tok.clone()
            vis_tok = O.erase_codebook_face(spec, vis_tok, **vc)
            assert torch.equal(vis_tok, ref_tok), (name, "erase_codebook_face")
        out["visual_tokens"] = vis_tok.clone()
    control_or = O.bert_control_emb(spec, sd, text, vis_tok)
    e = relerr(control_or, control_ref)
    assert e < 1e-6, (name, "control", e)
    out["control_emb"] = control_ref.clone() if control_ref.numel() < 2_000_000 else None
    # one forward over partially masked targets
    tgt = synth.synth_codes(B, spec.target_seq_len, seed=cfg["seed"] + 3)
    tgt[:, ::3] = spec.MASK
    emb = model.image_emb(tgt)
    tokens = torch.cat((control_ref, emb + model.target_pos_emb(emb)), dim=1)
    hid_ref = model.transformer_forward(tokens)
    logits_ref = model.to_logits(hid_ref[:, control_ref.shape[1]:])
    rel_ref = model.to_logits_rel(hid_ref[:, model.rel_tok_index])
    vid_ref = model.to_logits_vid(hid_ref[:, model.vid_tok_index])
    logits_or = torch.cat([O.bert_logits(spec, sd, control_or[i:i + 1], tgt[i:i + 1]) for i in range(B)], 0)
    e = relerr(logits_or, logits_ref)
    assert e < 5e-6, (name, "logits", e)
    out["oracle_logits_relerr"] = e
    out["target_in"] = tgt
    stride = 1 if logits_ref.numel() < 300_000 else 8
    out["logits_stride"] = stride
    out["logits"] = logits_ref[:, ::stride].clone()
    out["logits_argmax"] = logits_ref.argmax(-1).to(torch.int16)
    out["logits_norm"] = float(logits_ref.norm())
    out["rel_logit"], out["vid_logit"] = rel_ref.clone(), vid_ref.clone()
    print(f"{name}: control/logits oracle relerr ok ({e:.2e}); fwd {time.time() - t0:.1f}s")
    if cfg["dim"] <= 128:
        # full generate_images on CPU RNG (pins the sampler restatement; CUDA RNG differs by design)
        for dyn in (False, True):
            torch.manual_seed(cfg["seed"])
            images_ref, _, seq_ref = model.generate_images(text, visual=visual, mask_predict_steps=6,
                                                           mp_config=dict(O.DEFAULT_MP_CONFIG), dynamic=dyn)
            torch.manual_seed(cfg["seed"])
            images_or, seq_or = O.bert_generate_images(spec, sd, text, visual, steps=6, dynamic=dyn)
            assert torch.equal(seq_or, seq_ref), (name, "mask_predict ids", dyn)
            assert relerr(images_or, images_ref) < 1e-5
            out[f"gen_seq_dyn{int(dyn)}"] = seq_ref.clone()
            out[f"gen_images_dyn{int(dyn)}"] = images_ref.clone()
        # beam > 1 and preserve ('long' continuation, t_overlap=1)
        mpc = dict(O.DEFAULT_MP_CONFIG, B=2)
        torch.manual_seed(cfg["seed"] + 1)
        _, _, seq_ref = model.generate_images(text, visual=visual, mask_predict_steps=5, mp_config=mpc, dynamic=False,
                                              preserve=out["gen_seq_dyn0"], t_overlap=1, long_mode="long")
        torch.manual_seed(cfg["seed"] + 1)
        _, seq_or = O.bert_generate_images(spec, sd, text, visual, steps=5, mp_config=mpc, dynamic=False,
                                           preserve=out["gen_seq_dyn0"], t_overlap=1, long_mode="long")
        assert torch.equal(seq_or, seq_ref), (name, "beam/preserve ids")
        out["gen_seq_beam_preserve"] = seq_ref.clone()
        print(f"{name}: generate_images ids exact (dynamic on/off, beam 2 + preserve)")
    torch.save(out, fixture_path(name))


def gen_artv(name, cfg):
    ref_shims.install()
    from mmvid_pytorch.dalle_artv import DALLE
    vae, _ = build_ref_vae(cfg["image_size"], cfg["seed"] + 1000)
    cvae = build_ref_vae(cfg["image_size"], cfg["seed"] + 2000)[0]
    clip_sd = synth.clip_checkpoint_state_dict(cfg["dim"], cfg["layers"], seed=cfg["seed"])
    with ref_shims.fake_clip_checkpoint(clip_sd):
        model = DALLE(dim=cfg["dim"], vae=vae, cvae=cvae, num_text_tokens=cfg["vocab"],
                      text_seq_len=cfg["text_seq_len"], which_transformer="openai_clip_visual",
                      num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"], openai_clip_path="none")
    sd = synth.fill_state_dict(model, cfg["seed"], codebook_std=codebook_std())
    for k, v in vae.state_dict().items():
        sd["vae." + k] = v.clone()
    for k, v in cvae.state_dict().items():
        sd["cvae." + k] = v.clone()
    model.load_state_dict(sd, strict=True)
    model.eval()
    spec = O.ArtvSpec(dim=cfg["dim"], text_seq_len=cfg["text_seq_len"], num_text_tokens=cfg["vocab"],
                      num_visuals=cfg["num_visuals"], num_targets=cfg["num_targets"], image_size=cfg["image_size"])
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"])
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5)
    vis_tok = O.vae_get_codebook_indices(visual.reshape(-1, *visual.shape[2:]), O.sub_state_dict(sd, "cvae.")).view(B, -1)
    img_tok = synth.synth_codes(B, spec.target_seq_len, seed=cfg["seed"] + 3)
    logits_ref = model(text, visual=visual, target=img_tok)
    logits_or = O.artv_forward(spec, sd, text, vis_tok, img_tok)
    fin = torch.isfinite(logits_ref) & (logits_ref > -1e30)
    e = relerr(logits_or[fin], logits_ref[fin])
    assert e < 5e-6 and torch.equal(logits_or > -1e30, logits_ref > -1e30), (name, e)
    torch.manual_seed(cfg["seed"])
    images_ref, _, _ = model.generate_images(text, visual=visual)
    torch.manual_seed(cfg["seed"])
    seq_or = O.artv_generate_tokens(spec, sd, text, vis_tok)
    images_or = O.vae_decode(seq_or.reshape(-1, spec.image_seq_len), O.sub_state_dict(sd, "vae."))
    e2 = relerr(images_or.view_as(images_ref), images_ref)
    assert e2 < 1e-5, (name, "generate", e2)
    # image-token slice of the logits (the only unmasked columns in the image phase)
    lo = spec.num_control_tokens
    torch.save(dict(cfg=cfg, visual_tokens=vis_tok, image_tokens=img_tok,
                    image_logits=logits_ref[:, spec.control_seq_len:, lo:lo + spec.num_image_tokens].clone(),
                    gen_seq=seq_or.clone(), gen_images=images_ref.clone()), fixture_path(name))
    print(f"{name}: logits relerr {e:.2e}; generate_images (no-cache reference) matches oracle ({e2:.1e})")


def main():
    want = set(sys.argv[1:])
    torch.set_num_threads(os.cpu_count())
    for name, cfg in TRANSFORMER_CASES.items():
        if not want or name in want:
            gen_transformer(name, cfg)
    for name, cfg in VAE_CASES.items():
        if not want or name in want:
            gen_vae(name, cfg)
    for name, cfg in BERT_CASES.items():
        if not want or name in want:
            gen_bert(name, cfg)
    for name, cfg in ARTV_CASES.items():
        if not want or name in want:
            gen_artv(name, cfg)


if __name__ == "__main__":
    main()
