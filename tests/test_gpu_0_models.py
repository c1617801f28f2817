"""GPU parity of the model-level API against (a) the committed reference outputs (tests/golden/*.pt, produced
by the unmodified reference on CPU fp32) and (b) the CPU oracle run on the same synthetic weights.  The
sampled-id tests run the oracle ON THE GPU (torch eager fp32, TF32 off) because CPU and CUDA Philox streams
differ: same seed + same RNG call order => ids must be bit-identical in fp32 precision mode."""
import pytest
import torch

from cases import ARTV_CASES, BERT_CASES, TRANSFORMER_CASES, VAE_CASES
from helpers import (artv_spec, bert_spec, build_artv, build_bert, build_vae, load_fixture, relerr, to_device)
from mmvid_b200 import _lib, synth

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "tf32": 1e-3, "fp16": 1e-3, "bf16": 3e-2}


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


# ------------------------------------------------------------------------------------------------ transformer
@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
@pytest.mark.parametrize("name", list(TRANSFORMER_CASES))
def test_transformer_vs_reference_golden(name, prec):
    from mmvid_b200.transformer import OpenAICLIPTransformer
    cfg = TRANSFORMER_CASES[name]
    fx = load_fixture(name)
    m = OpenAICLIPTransformer(cfg["seq"], "openai_clip_visual", model_path=None, causal=True, mask_type=cfg["mask"],
                              mask_kwargs={"index": list(cfg["index"])}, width=cfg["dim"], layers=cfg["layers"],
                              precision=prec)
    m.load_state_dict(synth.fill_state_dict(m, cfg["seed"]))
    m = m.cuda().eval()
    x = synth.synth_tensor("x", (cfg["batch"], cfg["seq"], cfg["dim"]), cfg["seed"] + 7).cuda()
    x0 = x.clone()
    y = m(x)
    assert torch.equal(x, x0), "input must not be modified"
    e = relerr(y, fx["y"])
    print(f"{name} {prec}: relerr vs reference {e:.3e}")
    assert e < TOL[prec]


# ------------------------------------------------------------------------------------------------ VQGAN
@pytest.mark.parametrize("name", list(VAE_CASES))
def test_vae_indices_bit_exact_and_pixels(name):
    cfg = VAE_CASES[name]
    fx = load_fixture(name)
    vae, _ = build_vae(cfg["image_size"], cfg["seed"])
    img = synth.synth_frames(cfg["batch"], 1, cfg["image_size"], cfg["seed"])[:, 0].cuda()
    z = vae._encode_prequant(img)  # NHWC
    ez = relerr(z.permute(0, 3, 1, 2), fx["z"])
    idx = vae.get_codebook_indices(img)
    assert idx.dtype == torch.int64 and idx.shape == fx["indices"].shape
    n_diff = int((idx.cpu() != fx["indices"]).sum())
    print(f"{name}: pre-quant z relerr {ez:.2e}; index mismatches {n_diff}/{idx.numel()} (min top-2 gap {fx['min_top2_gap']:.2e})")
    assert ez < 1e-4
    assert n_diff == 0, "VQ codebook indices must be bit-exact"
    dec = vae.decode(fx["indices"].cuda())
    ed = relerr(dec, fx["decoded"])
    dec_r = vae.decode(fx["rand_codes"].cuda())
    er = relerr(dec_r, fx["decoded_rand"])
    print(f"{name}: decode relerr {ed:.2e} / random codes {er:.2e}")
    assert ed < 1e-3 and er < 1e-3
    assert float(dec.min()) >= 0.0 and float(dec.max()) <= 1.0


@pytest.mark.parametrize("name", ["vae_64", "vae_128"])
def test_vae_decode_tf32_tensor_core_pixels(name):
    """Decoder on tcgen05 (kind::tf32 convs + 1x1 GEMMs): pixels within 1e-3 of the reference's fp32 output."""
    cfg = VAE_CASES[name]
    fx = load_fixture(name)
    vae, _ = build_vae(cfg["image_size"], cfg["seed"], precision="tf32")
    dec = vae.decode(fx["indices"].cuda())
    e = relerr(dec, fx["decoded"])
    print(f"{name}: tf32 decode relerr {e:.2e}")
    assert e < 1e-3


@pytest.mark.parametrize("name", ["vae_64", "vae_128"])
def test_vae_decode_fp16_tensor_core_pixels_and_encoder_untouched(name):
    """precision='fp16': the decoder's 3x3 convs run kind::f16 on fp16 activations; pixels within 1e-3 of the reference's
    fp32 output.  The ENCODER of the same module stays on the tf32 kernels (its indices must not depend on the decoder's
    throughput mode): the mismatch rate of its indices against the reference's is reported, next to the tf32 module's."""
    cfg = VAE_CASES[name]
    fx = load_fixture(name)
    vae, _ = build_vae(cfg["image_size"], cfg["seed"], precision="fp16")
    dec = vae.decode(fx["indices"].cuda())
    e = relerr(dec, fx["decoded"])
    img = synth.synth_frames(cfg["batch"], 1, cfg["image_size"], cfg["seed"])[:, 0].cuda()
    idx16 = vae.get_codebook_indices(img).cpu()
    vae32, _ = build_vae(cfg["image_size"], cfg["seed"], precision="tf32")
    idx32 = vae32.get_codebook_indices(img).cpu()
    mm16, mm32 = int((idx16 != fx["indices"]).sum()), int((idx32 != fx["indices"]).sum())
    print(f"{name}: fp16 decode relerr {e:.2e}; encoder index mismatches vs reference: {mm16}/{idx16.numel()} (fp16 module), "
          f"{mm32}/{idx32.numel()} (tf32 module) - bit-exact indices are the fp32 module's contract")
    assert e < 1e-3
    assert torch.equal(idx16, idx32)


@pytest.mark.parametrize("prec", ["tf32", "fp16"])
def test_vae_decode_with_conv_fused_groupnorm_statistics_equals_the_two_pass_path(monkeypatch, prec):
    """Decoder convs write the GroupNorm partial statistics of their results (MMVID_GN_FUSE=1, default): same pixels as with
    the separate statistics pass (MMVID_GN_FUSE=0) up to the summation order of the statistics."""
    cfg = VAE_CASES["vae_128"]
    fx = load_fixture("vae_128")
    vae, _ = build_vae(cfg["image_size"], cfg["seed"], precision=prec)
    ids = fx["indices"].cuda()
    monkeypatch.setenv("MMVID_GN_FUSE", "0")
    two_pass = vae.decode(ids)
    monkeypatch.setenv("MMVID_GN_FUSE", "1")
    n0 = _lib.launch_count()
    fused = vae.decode(ids)
    n_fused = _lib.launch_count() - n0
    monkeypatch.setenv("MMVID_GN_FUSE", "0")
    n0 = _lib.launch_count()
    vae.decode(ids)
    n_two = _lib.launch_count() - n0
    e = relerr(fused, two_pass)
    print(f"vae_128 {prec}: fused-statistics decode vs two-pass decode relerr {e:.2e}; launches {n_fused} vs {n_two}; "
          f"vs reference {relerr(fused, fx['decoded']):.2e}")
    # (statistics that agree to ~1e-6 still move some activations across a 10-bit rounding boundary of the next conv's
    # operands: the two decodes are two realisations of the same rounding noise, each equally far from the reference)
    assert e < 1e-3 and relerr(fused, fx["decoded"]) < 1e-3
    assert n_fused < n_two  # statistics kernels actually left the launch list


# ------------------------------------------------------------------------------------------------ BERT
_BERT_CACHE = {}


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["bert_tiny", "bert_tiny_nov", "bert_shapeB", "bert_shapeB_vis", "bert_shapeA"])
def test_bert_forward_vs_reference_golden(name, prec):
    cfg = BERT_CASES[name]
    if prec not in ("tf32", "fp16") and name == "bert_shapeA":
        pytest.skip("Shape A is checked in the two <= 1e-3 tensor-core modes only (the fp32 CUDA-core path is slow)")
    fx = load_fixture(name)
    if name not in _BERT_CACHE:
        _BERT_CACHE.clear()  # keep at most one big model resident
        _BERT_CACHE[name] = build_bert(cfg, precision=prec)[0]
    model = _BERT_CACHE[name]
    model.precision = prec
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = None
    if cfg["num_visuals"] > 0:
        visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
        vt = model.get_image_tokens(visual, which_vae="cvae")
        raw = fx.get("visual_tokens_raw", fx["visual_tokens"])
        assert torch.equal(vt.cpu(), raw), "visual-control VQ ids must be bit-exact"
    # BASELINE config 4: the visual-control grid windowed by the reference's vc_mode hook (dalle_bert.py:950-953)
    vc = dict(vc_mode=cfg["vc_mode"], face_mode=cfg.get("face_mode")) if cfg.get("vc_mode") else {}
    if vc:
        assert torch.equal(model.erase_codebook_face(vt.clone(), **vc).cpu(), fx["visual_tokens"])
    control = model(text, visual=visual, return_loss=False, **vc)
    if fx["control_emb"] is not None:
        assert relerr(control, fx["control_emb"]) < 1e-6
    tgt = fx["target_in"].cuda()
    x = torch.empty(B, model.total_seq_len, cfg["dim"], device="cuda")
    x[:, : control.shape[1]] = control
    from mmvid_b200 import ops
    ops.embed_gather(x, [model._target_segment(tgt)])
    hid = model.transformer_forward(x)
    csl = control.shape[1]
    logits = model._head(hid[:, csl:].reshape(-1, cfg["dim"]), model.to_logits).view(B, -1, 1024)
    st = fx["logits_stride"]
    e = relerr(logits[:, ::st], fx["logits"])
    agree = float((logits.argmax(-1).cpu() == fx["logits_argmax"].long()).float().mean())
    rel = model._head_scalar(hid[:, model.rel_tok_index].contiguous(), model.to_logits_rel)
    vid = model._head_scalar(hid[:, model.vid_tok_index].contiguous(), model.to_logits_vid)
    e_rel = float((rel.cpu() - fx["rel_logit"]).abs().max())
    e_vid = float((vid.cpu() - fx["vid_logit"]).abs().max())
    print(f"{name} {prec}: logits relerr {e:.3e}, argmax agreement {agree:.4f}, |d rel| {e_rel:.2e}, |d vid| {e_vid:.2e}")
    assert e < TOL[prec]
    if prec != "bf16":
        assert e_rel < 5e-3 and e_vid < 5e-3


def test_bert_shapeB_sampled_ids_bit_exact_vs_oracle_on_gpu_full_width():
    """Sampled ids at the reference scripts' own model size (768 x 12, Shape B, text + visual control through the cVAE):
    fp32 CUDA-core mode, reference RNG order, against the oracle's sampler running on the same GPU under the same seed."""
    from oracle import mmvid_oracle as O
    cfg = dict(BERT_CASES["bert_shapeB_vis"], batch=2)
    _BERT_CACHE.clear()
    model, sd = build_bert(cfg, precision="fp32")
    sd_dev = to_device(sd, "cuda")
    spec = bert_spec(cfg)
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
    for dyn, steps in ((False, 3), (True, 4)):
        torch.manual_seed(321)
        images, _, seq = model.generate_images(text, visual=visual, mask_predict_steps=steps, dynamic=dyn)
        torch.manual_seed(321)
        images_o, seq_o = O.bert_generate_images(spec, sd_dev, text, visual, steps=steps, dynamic=dyn)
        n_diff = int((seq != seq_o).sum())
        print(f"bert_shapeB_vis fp32 ids (dynamic={dyn}, steps={steps}): {n_diff} of {seq.numel()} differ; "
              f"frames relerr {relerr(images, images_o):.2e}")
        assert n_diff == 0
        assert relerr(images, images_o) < 1e-4
    del model, sd_dev
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["bert_tiny", "bert_tiny_nov"])
def test_bert_generate_images_ids_bit_exact_vs_oracle_on_gpu(name):
    from oracle import mmvid_oracle as O
    cfg = BERT_CASES[name]
    model, sd = build_bert(cfg, precision="fp32")
    sd_dev = to_device(sd, "cuda")
    spec = bert_spec(cfg)
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = None
    if cfg["num_visuals"] > 0:
        visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
    for dyn, mpc, steps in ((False, None, 6), (True, None, 8), (False, dict(O.DEFAULT_MP_CONFIG, B=2), 5)):
        torch.manual_seed(123)
        images, _, seq = model.generate_images(text, visual=visual, mask_predict_steps=steps, dynamic=dyn, mp_config=mpc)
        torch.manual_seed(123)
        images_o, seq_o = O.bert_generate_images(spec, sd_dev, text, visual, steps=steps, dynamic=dyn, mp_config=mpc)
        assert torch.equal(seq, seq_o), f"{name}: sampled ids differ (dynamic={dyn}, mp={mpc is not None})"
        assert images.shape == images_o.shape and relerr(images, images_o) < 1e-4
    # preserve / long-video continuation
    torch.manual_seed(5)
    _, _, seq2 = model.generate_images(text, visual=visual, mask_predict_steps=4, dynamic=False, preserve=seq, t_overlap=1)
    torch.manual_seed(5)
    _, seq2_o = O.bert_generate_images(spec, sd_dev, text, visual, steps=4, dynamic=False, preserve=seq_o, t_overlap=1)
    assert torch.equal(seq2, seq2_o)
    n = spec.image_seq_len
    assert torch.equal(seq2.view(B, -1)[:, :n], seq.view(B, -1)[:, -n:])
    # temporal interpolation (visualize_long, utils_train.py:1378-1431): the first half of the previous codes becomes every
    # other frame of the next clip, the frames in between are predicted
    if spec.num_targets % 2 == 0:
        Ttot = spec.target_seq_len
        prev = seq.view(B, Ttot)
        preserve = torch.full_like(prev, spec.MASK)
        preserve[:, : Ttot // 2] = prev[:, : Ttot // 2]
        torch.manual_seed(6)
        _, _, seq3 = model.generate_images(text, visual=visual, mask_predict_steps=4, dynamic=False, preserve=preserve,
                                           t_overlap=1, long_mode="interp")
        torch.manual_seed(6)
        _, seq3_o = O.bert_generate_images(spec, sd_dev, text, visual, steps=4, dynamic=False, preserve=preserve, t_overlap=1,
                                           long_mode="interp")
        assert torch.equal(seq3, seq3_o)
        kept = seq3.view(B, spec.num_targets, n)[:, ::2]
        assert torch.equal(kept, prev.view(B, spec.num_targets, n)[:, : spec.num_targets // 2])


def test_bert_batched_sampler_is_valid_and_seeded():
    cfg = BERT_CASES["bert_tiny_nov"]
    model, _ = build_bert(cfg, precision="tf32", sampling_mode="batched")
    text = synth.synth_text(4, cfg["text_seq_len"], cfg["vocab"], 3).cuda()
    torch.manual_seed(1)
    im1, _, s1 = model.generate_images(text, mask_predict_steps=5, dynamic=False)
    torch.manual_seed(1)
    im2, _, s2 = model.generate_images(text, mask_predict_steps=5, dynamic=False)
    assert torch.equal(s1, s2) and torch.equal(im1, im2)
    assert s1.min() >= 0 and s1.max() < 1024
    assert im1.shape == (4, cfg["num_targets"], 3, cfg["image_size"], cfg["image_size"])


@pytest.mark.parametrize("beams,dynamic", [(1, False), (1, True), (3, False), (2, True)])
def test_device_resident_mask_predict_beams_and_dynamic_stop(beams, dynamic, monkeypatch):
    """sampling_mode='batched' = the device-resident loop (fused draws, beam scoring / choice and dynamic stop on the device,
    no host sync): valid ids, reproducible under torch.manual_seed, CUDA-graph replay == eager launches, and for beam 1
    without dynamic stop statistically the same sampler as the torch-op loop (mean log-probability of the final ids under a
    common scoring forward)."""
    cfg = BERT_CASES["bert_tiny_nov"]
    model, _ = build_bert(cfg, precision="tf32", sampling_mode="batched")
    text = synth.synth_text(4, cfg["text_seq_len"], cfg["vocab"], 3).cuda()
    mpc = dict(T1_n=10, T2_n=10, T3_n=30, N1_n=0.9, N2_n=0.1, N3_n=0.125, N4_n=0.0625, T1_t=10, T2_t=5, T3_t=35, N1_t=0.0,
               N2_t=0.0, N3_t=0.0, N4_t=0.0, T=8, B=beams)
    outs = {}
    for sw in ("1", "0", "1"):
        monkeypatch.setenv("MMVID_CUDA_GRAPH", sw)
        torch.manual_seed(21)
        im, _, s = model.generate_images(text, mask_predict_steps=8, dynamic=dynamic, mp_config=mpc)
        assert s.min() >= 0 and s.max() < 1024 and im.shape[0] == 4
        if sw in outs:
            assert torch.equal(outs[sw], s)
        outs[sw] = s
    assert torch.equal(outs["0"], outs["1"])
    torch.manual_seed(22)
    _, _, s2 = model.generate_images(text, mask_predict_steps=8, dynamic=dynamic, mp_config=mpc)
    assert not torch.equal(s2, outs["1"])
    if beams == 1 and not dynamic:
        # same sampler as the torch-op loop: compare how likely each loop's final ids are under the model itself
        def mean_logp(seq):
            control = model(text.repeat(seq.shape[0] // (4 * cfg["num_targets"]), 1), return_loss=False)
            ids = seq.view(control.shape[0], -1)
            x = torch.empty(control.shape[0], model.total_seq_len, cfg["dim"], device="cuda")
            x[:, :control.shape[1]] = control
            from mmvid_b200 import ops
            ops.embed_gather(x, [model._target_segment(ids)])
            hid = model.transformer_forward(x)
            lg = model._head(hid[:, control.shape[1]:].reshape(-1, cfg["dim"]), model.to_logits).view(control.shape[0], -1, 1024)
            return float(torch.log_softmax(lg, -1).gather(2, ids.unsqueeze(-1)).mean())
        a, b = [], []
        for rep in range(6):
            monkeypatch.setenv("MMVID_SAMPLER", "device")
            torch.manual_seed(100 + rep)
            a.append(mean_logp(model.generate_images(text, mask_predict_steps=8, dynamic=False, mp_config=mpc)[2]))
            monkeypatch.setenv("MMVID_SAMPLER", "torch")
            torch.manual_seed(100 + rep)
            b.append(mean_logp(model.generate_images(text, mask_predict_steps=8, dynamic=False, mp_config=mpc)[2]))
        ma, mb = sum(a) / len(a), sum(b) / len(b)
        print(f"mean log-prob of final ids: device sampler {ma:.4f}, torch-op sampler {mb:.4f}")
        assert abs(ma - mb) < 0.15 * max(1.0, abs(mb))


# ------------------------------------------------------------------------------------------------ ART-V
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_artv_forward_logits_vs_reference_golden(prec):
    cfg = ARTV_CASES["artv_tiny"]
    fx = load_fixture("artv_tiny")
    model, _ = build_artv(cfg, precision=prec)
    spec = artv_spec(cfg)
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
    logits = model(text, visual=visual, target=fx["image_tokens"].cuda())
    assert logits.shape == (B, spec.total_seq_len, spec.total_tokens)
    lo = spec.num_control_tokens
    e = relerr(logits[:, spec.control_seq_len:, lo:], fx["image_logits"])
    print(f"artv_tiny {prec}: image-logit relerr {e:.3e}")
    assert e < TOL[prec]
    assert float(logits[:, spec.control_seq_len:, :lo].max()) < -1e30  # text/visual columns masked in image rows
    assert float(logits[:, :spec.text_seq_len, spec.num_text_tokens:].max()) < -1e30


@pytest.mark.parametrize("reps,impl", [(1, "fused"), (4, "fused"), (1, "native"), (1, "persistent"), (5, "native"), (9, "native")])
def test_artv_kv_cache_generate_matches_no_cache_oracle_on_gpu(reps, impl):
    """B=2 through the native per-layer launches and through the persistent cooperative decode kernel; reps=5 -> B=10:
    native; reps=9 -> B=18: generic path.  All must reproduce the reference's full re-forward sampling bit for bit."""
    from oracle import mmvid_oracle as O
    cfg = ARTV_CASES["artv_tiny"]
    model, sd = build_artv(cfg, precision="fp32")
    model.decode_impl = impl
    sd_dev = to_device(sd, "cuda")
    spec = artv_spec(cfg)
    B = cfg["batch"] * reps
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
    vis_tok = model.get_image_tokens(visual, which_vae="cvae")
    torch.manual_seed(77)
    images, _, toks = model.generate_images(text, visual=visual, return_tokens=True)
    torch.manual_seed(77)
    toks_o = O.artv_generate_tokens(spec, sd_dev, text, vis_tok)
    assert torch.equal(toks, toks_o), "KV-cache decode must reproduce the reference's full re-forward sampling"
    fx = load_fixture("artv_tiny")
    assert images.shape[1:] == fx["gen_images"].shape[1:] and images.shape[0] == B


@pytest.mark.parametrize("prec", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("reps", [1, 2, 4])
def test_artv_streaming_decode_tiny_matches_fp32_kv_cache_path(prec, reps):
    """decode_stream.cu on the tiny model (D = 128: most CTAs own no column of the narrow matrices, Z = 1 split), B = 2 / 4 /
    8: per-step image logits against the fp32 fused path fed with the SAME sampled tokens."""
    cfg = ARTV_CASES["artv_tiny"]
    B = cfg["batch"] * reps
    model, _ = build_artv(cfg, precision=prec, sampling_mode="batched")
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
    trace = []
    torch.manual_seed(3)
    _, _, toks = model.generate_images(text, visual=visual, return_tokens=True, logits_trace=trace)
    step = torch.stack(trace, 1)
    model.precision = model.transformer.precision = "fp32"
    spec = artv_spec(cfg)
    full = model(text, visual=visual, target=toks)
    P = spec.control_seq_len + 1  # <bos> + text + visual tokens: row P - 1 predicts the first image token
    lo = model.num_control_tokens
    rows = full[:, P - 1:P - 1 + toks.shape[1], lo:lo + 1024]
    e = relerr(step, rows)
    print(f"artv_tiny streaming decode {prec} B={B}: logits relerr vs fp32 full forward {e:.2e}")
    assert e < (3e-2 if prec == "bf16" else 2e-3)


def test_artv_shapeB_kv_cache_first_tokens_match_no_cache_oracle_full_width():
    """BASELINE config 3 at the reference scripts' size (768 x 12, Shape B, batch 4): the first 32 sampled tokens of the
    KV-cache decode against the oracle's full re-forward sampling on the same GPU (fp32 mode, same seed)."""
    from oracle import mmvid_oracle as O
    cfg = dict(dim=768, layers=12, text_seq_len=50, vocab=49408, num_visuals=1, num_targets=8, image_size=128, seed=33,
               batch=4)
    model, sd = build_artv(cfg, precision="fp32")
    sd_dev = to_device(sd, "cuda")
    spec = artv_spec(cfg)
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], cfg["seed"] + 5).cuda()
    vis_tok = model.get_image_tokens(visual, which_vae="cvae")
    torch.manual_seed(78)
    toks = model.generate_tokens(text, visual=visual, max_new=32)
    torch.manual_seed(78)
    toks_o = O.artv_generate_tokens(spec, sd_dev, text, vis_tok, max_new=32)
    n_diff = int((toks != toks_o).sum())
    print(f"artv shape B, batch 4: {n_diff} of {toks.numel()} sampled ids differ over the first 32 steps")
    assert toks.shape == toks_o.shape == (B, 32) and n_diff == 0
    del model, sd_dev
    torch.cuda.empty_cache()


def test_batched_mask_predict_cuda_graph_replay_equals_eager_launches(monkeypatch):
    """The forward chain captured in a CUDA graph (static ids / x buffers) must produce exactly the ids and frames of the
    launch-by-launch path under the same seed, also after the weights change (cache key = parameter versions)."""
    cfg = BERT_CASES["bert_tiny_nov"]
    model, _ = build_bert(cfg, precision="tf32", sampling_mode="batched")
    text = synth.synth_text(4, cfg["text_seq_len"], cfg["vocab"], 3).cuda()
    outs = {}
    for sw in ("0", "1", "1"):
        monkeypatch.setenv("MMVID_CUDA_GRAPH", sw)
        torch.manual_seed(7)
        im, _, s = model.generate_images(text, mask_predict_steps=5, dynamic=False)
        if sw in outs:
            assert torch.equal(outs[sw][1], s)  # second graphed call: replay of the cached graph
        outs[sw] = (im, s)
    assert torch.equal(outs["0"][1], outs["1"][1]) and torch.equal(outs["0"][0], outs["1"][0])
    with torch.no_grad():
        model.to_logits[1].bias.add_(0.5 * torch.randn_like(model.to_logits[1].bias))
    res = {}
    for sw in ("1", "0"):
        monkeypatch.setenv("MMVID_CUDA_GRAPH", sw)
        torch.manual_seed(7)
        _, _, res[sw] = model.generate_images(text, mask_predict_steps=5, dynamic=False)
    assert torch.equal(res["0"], res["1"]) and not torch.equal(res["1"], outs["1"][1])
