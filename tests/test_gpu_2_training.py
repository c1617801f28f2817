"""GPU: BERT.forward(return_loss=True) - loss values and parameter gradients of the CUDA training path against
the oracle's torch-autograd restatement of dalle_bert.py:1030-1127 on the same weights, masks and negatives."""
import numpy as np
import pytest
import torch

from cases import BERT_CASES
from helpers import bert_spec, build_bert, relerr, to_device
from mmvid_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference_math_with_grad():
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.enable_grad():
        yield


@pytest.mark.parametrize("batched", [True, False], ids=["one_pass_3B", "three_passes"])
@pytest.mark.parametrize("prec,tol_loss,tol_grad", [("fp32", 2e-5, 2e-4), ("tf32", 2e-3, 2e-2)])
@pytest.mark.parametrize("name", ["bert_tiny", "bert_tiny_nov"])
def test_training_losses_and_gradients_match_oracle_autograd(name, prec, tol_loss, tol_grad, batched):
    """batched: the positive pass and the REL / VID negatives as ONE transformer pass over a 3B batch (default) vs the
    reference's three passes (dalle_bert.py:1037,1057,1101) - both against the oracle's three-pass autograd."""
    from oracle import mmvid_oracle as O
    cfg = BERT_CASES[name]
    model, sd = build_bert(cfg, precision=prec)
    model.batch_train_passes = batched
    model.train()
    spec = bert_spec(cfg)
    B = cfg["batch"]
    g = torch.Generator().manual_seed(3)
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
    vis_ids = torch.randint(0, 1024, (B, spec.visual_seq_len), generator=g).cuda() if cfg["num_visuals"] > 0 else None
    tgt = torch.randint(0, 1024, (B, spec.target_seq_len), generator=g).cuda()
    warp = torch.randint(0, 1024, (B, spec.target_seq_len), generator=g).cuda()
    mask1 = (torch.rand(B, spec.target_seq_len, generator=g) < 0.4).cuda()
    mask1[0, :2] = False
    nfm = torch.ones(B).cuda()
    # oracle (torch autograd, fp32, same GPU)
    sd_o = {k: v.clone().cuda().requires_grad_(v.is_floating_point() and not k.startswith(("vae.", "cvae.")))
            for k, v in sd.items()}
    lo = O.bert_train_losses(spec, sd_o, text, vis_ids, tgt, mask1, nfm, rel=True, vid=True, target_warp_tokens=warp)
    (7 * lo[0] + 0.5 * lo[1] + 0.5 * lo[2]).backward()   # train.py:320 weighting (utils_args.py:399-410)
    lm = model._losses(text, vis_ids, tgt, mask1, nfm, rel=True, vid=True, target_warp_ids=warp)
    (7 * lm[0] + 0.5 * lm[1] + 0.5 * lm[2]).backward()
    for a, b, nm in zip(lm, lo, ("msm", "rel", "vid")):
        e = abs(float(a) - float(b)) / max(abs(float(b)), 1e-6)
        print(f"{name} {prec} loss_{nm}: {float(a):.6f} vs oracle {float(b):.6f} (rel {e:.2e})")
        assert e < tol_loss
    checked = 0
    worst = 0.0
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        go = sd_o[k].grad
        assert p.grad is not None, f"no gradient for {k}"
        if go is None or float(go.norm()) == 0:
            continue
        e = relerr(p.grad, go)
        worst = max(worst, e)
        assert e < tol_grad, f"{k}: grad relerr {e:.3e}"
        checked += 1
    print(f"{name} {prec}: {checked} parameter gradients checked, worst relerr {worst:.2e}")
    assert checked > 20


@pytest.mark.parametrize("fused", [False, True])
def test_reference_train_call_signature_and_optimizer_step(fused):
    """The call train.py:298-325 makes: losses -> weighted sum -> backward -> clip -> Adam step; loss must go down.
    fused=False: torch's own Adam / clip on the drop-in module; fused=True: the library's multi-tensor kernels."""
    from mmvid_b200 import optim as FO
    cfg = BERT_CASES["bert_tiny"]
    model, _ = build_bert(cfg, precision="tf32")
    model.train()
    B = 4
    np.random.seed(0)
    torch.manual_seed(0)
    import random
    random.seed(0)
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], 1).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], 2).cuda()
    frames = synth.synth_frames(B, cfg["num_targets"], cfg["image_size"], 3).cuda()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = FO.FusedAdam(params, lr=3e-3) if fused else torch.optim.Adam(params, lr=3e-3)
    hist = []
    for it in range(6):
        np.random.seed(1)
        random.seed(1)
        torch.manual_seed(1)  # same masks every iteration so the loss is comparable
        loss_msm, loss_rel, loss_vid = model(text, visual=visual, target=frames, return_loss=True, rel=True, vid=True,
                                             msm_strategy_prob=np.array([0.7, 0.1, 0.1, 0.1]),
                                             msm_bernoulli_prob=[0.2, 0.5], vid_strategy_prob=np.array([0.25] * 4))
        loss = 7 * loss_msm + 0.5 * loss_rel + 0.5 * loss_vid
        opt.zero_grad()
        loss.backward()
        if fused:
            FO.clip_grad_norm_(params, 1.0)
        else:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        hist.append(float(loss))
    print("training loss history:", [round(h, 4) for h in hist])
    assert all(np.isfinite(hist)) and hist[-1] < hist[0]


def test_shipped_training_flags_negvc_and_visual_aug_mode_run_and_match_oracle_losses():
    """scripts/mmvoxceleb/image_and_video/train.sh passes --visual_aug_mode motion_color; --negvc feeds an explicit negative
    text (train.py:259-262, 312-314).  negvc: the REL negative is [REL] + text_neg + [ST1][VID] WITHOUT a visual segment
    (dalle_bert.py:909-935, 974-975), checked here against a direct evaluation of that sequence."""
    import random
    cfg = BERT_CASES["bert_tiny"]
    model, _ = build_bert(cfg, precision="fp32")
    model.train()
    B = 2
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], 1).cuda()
    text_neg = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], 9).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], 2).cuda()
    frames = synth.synth_frames(B, cfg["num_targets"], cfg["image_size"], 3).cuda()
    kw = dict(target=frames, return_loss=True, rel=True, vid=True, msm_strategy_prob=np.array([1.0, 0, 0, 0]),
              msm_bernoulli_prob=[0.4, 0.6], vid_strategy_prob=np.array([0.25] * 4))

    def run(**extra):
        np.random.seed(3); random.seed(3); torch.manual_seed(3)
        return model(text, visual=visual, **kw, **extra)

    base = run()
    neg = run(negvc=True, text_neg=text_neg, visual_neg=visual)
    aug = run(visual_aug_mode="motion_color")
    for l in (*base, *neg, *aug):
        assert torch.isfinite(l)
    assert float(neg[0]) == float(base[0]) and float(neg[2]) == float(base[2])  # MSM / VID untouched by negvc
    assert float(neg[1]) != float(base[1])
    # direct evaluation of the reference's negative sequence: same masks (same seeds), REL logit of the shorter sequence
    np.random.seed(3); random.seed(3); torch.manual_seed(3)
    with torch.no_grad():
        vis_ids = model.get_image_tokens(visual, which_vae="cvae")
        tgt_ids = model.get_image_tokens(frames)
        mask1, _ = model._sample_msm_masks(B, text.device, kw["msm_strategy_prob"], kw["msm_bernoulli_prob"], 0)
        tgt_masked = torch.where(mask1, tgt_ids, torch.full_like(tgt_ids, 1024))
        control, temb = model._embed_train(text, vis_ids, tgt_masked)
        cneg = model._embed_train(text_neg, None, tgt_masked, with_visual=False, with_target=False)[0]
        assert cneg.shape[1] == 1 + cfg["text_seq_len"] + 2
        lp = model._head_train(model._transformer_train(torch.cat((control, temb), 1))[:, 0], model.to_logits_rel).squeeze(-1)
        ln = model._head_train(model._transformer_train(torch.cat((cneg, temb), 1))[:, 0], model.to_logits_rel).squeeze(-1)
        bce = torch.nn.functional.binary_cross_entropy_with_logits
        want = bce(lp, torch.ones(B, device="cuda")) + bce(ln, torch.zeros(B, device="cuda"))
    assert abs(float(neg[1]) - float(want)) < 1e-5 * max(1.0, abs(float(want)))
    (7 * neg[0] + 0.5 * neg[1] + 0.5 * neg[2]).backward()
    assert model.text_emb.weight.grad is not None and float(model.text_emb.weight.grad.abs().sum()) > 0


@pytest.mark.parametrize("kind,wd", [("adam", 0.0), ("adam", 0.01), ("adamw", 0.05)])
def test_fused_adam_and_clip_match_torch_optimisers(kind, wd):
    """train.py:322-325 step: clip_grad_norm_ + Adam/AdamW, multi-tensor kernels vs torch's own CPU optimisers
    (the reference's definition, utils_train.py:167-181)."""
    from mmvid_b200 import optim as FO
    g = torch.Generator().manual_seed(5)
    shapes = [(768, 768), (3, 5, 7), (1,), (70001,), (1026, 64), (131072,)]
    ref = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]
    kw = dict(lr=3e-3, weight_decay=wd)
    if kind == "adam":
        o_ref, o_ours = torch.optim.Adam(ref, **kw), FO.FusedAdam(ours, **kw)
    else:
        o_ref, o_ours = torch.optim.AdamW(ref, betas=(0.9, 0.95), **kw), FO.FusedAdamW(ours, betas=(0.9, 0.95), **kw)
    for it in range(4):
        o_ref.zero_grad()
        o_ours.zero_grad()
        for i, (a, b) in enumerate(zip(ref, ours)):
            if it == 2 and i == 1:
                continue  # a parameter without a gradient this step
            gr = torch.randn(a.shape, generator=g) * (10.0 if it % 2 == 0 else 0.01)  # clipped / not clipped
            a.grad = gr.clone()
            b.grad = gr.cuda()
        n_ref = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        n_ours = FO.clip_grad_norm_(ours, 1.0)
        assert abs(float(n_ours) - float(n_ref)) <= 1e-5 * float(n_ref)
        for a, b in zip(ref, ours):
            if a.grad is not None:
                assert relerr(b.grad.cpu(), a.grad) < 1e-5  # clip coefficient: sqrt(sum g^2) vs torch's norm of norms
        o_ref.step()
        o_ours.step()
        for a, b in zip(ref, ours):
            d = (b.detach().cpu() - a.detach()).abs()
            # m / (sqrt(v) + eps) is ill-conditioned where grad + wd * p cancels to ~eps (fma vs mul+add rounding decides
            # the sign): allow a vanishing fraction of such elements, bounded by one full update of size lr
            assert float(d.mean()) < 1e-8 and float((d > 2e-6).float().mean()) < 1e-4 and float(d.max()) <= 2 * 3e-3, \
                (it, a.shape, float(d.max()))
    sd = o_ours.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}  # torch's checkpoint layout (train.py:352)
    o_ref.load_state_dict({"state": {k: {kk: (vv.cpu() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
                                     for k, v in sd["state"].items()}, "param_groups": sd["param_groups"]})


def test_sampling_after_a_fused_adam_step_uses_the_updated_weights():
    """FusedAdam writes parameters through raw pointers (no torch version bump): the axial position tables, 16-bit weight
    copies and the captured CUDA graph must still follow.  Sample, take one large step, sample again, and compare with a
    FRESH model loaded from the trained state dict (train.py samples periodically during training)."""
    import random
    from mmvid_b200 import optim as FO
    cfg = BERT_CASES["bert_tiny"]
    model, _ = build_bert(cfg, precision="tf32", sampling_mode="batched")
    B = 2
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], 1).cuda()
    visual = synth.synth_frames(B, cfg["num_visuals"], cfg["image_size"], 2).cuda()
    frames = synth.synth_frames(B, cfg["num_targets"], cfg["image_size"], 3).cuda()

    def sample(m):
        torch.manual_seed(11)
        with torch.no_grad():
            return m.generate_images(text, visual=visual, mask_predict_steps=3, dynamic=False)[2]

    def logits_of(m):
        with torch.no_grad():
            control = m(text, visual=visual, return_loss=False)
            x = torch.zeros(B, m.total_seq_len, cfg["dim"], device="cuda")
            x[:, :control.shape[1]] = control
            from mmvid_b200 import ops
            ids = torch.full((B, m.target_seq_len), 7, dtype=torch.long, device="cuda")
            ops.embed_gather(x, [m._target_segment(ids)])
            return m.transformer_forward(x).clone()

    model.eval()
    seq0, h0 = sample(model), logits_of(model)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = FO.FusedAdam(params, lr=0.05)
    np.random.seed(1); random.seed(1); torch.manual_seed(1)
    losses = model(text, visual=visual, target=frames, return_loss=True, rel=True, vid=True,
                   msm_strategy_prob=np.array([0.7, 0.1, 0.1, 0.1]), msm_bernoulli_prob=[0.2, 0.5],
                   vid_strategy_prob=np.array([0.25] * 4))
    opt.zero_grad()
    (7 * losses[0] + 0.5 * losses[1] + 0.5 * losses[2]).backward()
    opt.step()
    model.eval()
    seq1, h1 = sample(model), logits_of(model)
    fresh, _ = build_bert(cfg, precision="tf32", sampling_mode="batched")
    fresh.load_state_dict({k: v.detach().clone() for k, v in model.state_dict().items()})
    fresh.eval()
    seq2, h2 = sample(fresh), logits_of(fresh)
    assert relerr(h1, h0) > 1e-2, "the step was too small to tell stale weights from fresh ones"
    assert relerr(h1, h2) < 1e-6, "hidden states after the step differ from a fresh model with the same weights"
    assert torch.equal(seq1, seq2)
    print(f"after one FusedAdam step: hidden relerr vs before {relerr(h1, h0):.2e}, vs fresh model {relerr(h1, h2):.1e}; "
          f"{int((seq1 != seq0).sum())} of {seq1.numel()} sampled ids changed")
