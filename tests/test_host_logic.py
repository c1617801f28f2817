"""CPU: host-side logic of the drop-in classes (no kernels launched): shapes, state-dict keys, schedules,
visual-control erasing, batch sharding arithmetic."""
import numpy as np
import pytest
import torch

from cases import BERT_CASES
from helpers import build_artv, build_bert
from oracle import mmvid_oracle as O


def test_bert_sequence_bookkeeping_matches_reference_formulas():
    cfg = dict(BERT_CASES["bert_tiny"])
    m, _ = build_bert(cfg, device="cpu")
    n = (cfg["image_size"] // 16) ** 2
    assert m.image_seq_len == n and m.target_seq_len == cfg["num_targets"] * n
    assert m.total_seq_len == 1 + cfg["text_seq_len"] + cfg["num_visuals"] * n + 2 + cfg["num_targets"] * n
    assert m.st1_tok_index == 1 + cfg["text_seq_len"] + cfg["num_visuals"] * n and m.vid_tok_index == m.st1_tok_index + 1
    assert m.transformer.mask_rows == (m.st1_tok_index, m.vid_tok_index)
    assert m.image_token_lut == {"[MASK]": 1024, "[SEP]": 1025}
    assert m.text_emb.weight.shape[0] == cfg["vocab"] + cfg["text_seq_len"]
    assert all(not p.requires_grad for p in m.vae.parameters())


def test_shape_A_and_B_sequence_lengths():
    # SURVEY.md appendix B shape calculator
    for L, px, V, want in ((64, 256, 0, 2115), (64, 256, 1, 2371), (50, 128, 0, 565), (50, 128, 1, 629)):
        s = O.BertSpec(dim=768, text_seq_len=L, num_text_tokens=49408, num_visuals=V, num_targets=8, image_size=px)
        assert s.total_seq_len == want
    a = O.ArtvSpec(dim=768, text_seq_len=64, num_text_tokens=49408, num_visuals=1, num_targets=8, image_size=256)
    assert (a.total_seq_len, a.total_tokens, a.num_control_tokens) == (2368, 51776, 50752)


def test_mask_predict_schedules_match_oracle():
    from mmvid_b200.dalle_bert import DEFAULT_MP_CONFIG, mask_predict_schedules
    for N in (8, 512, 2048, 1792):
        assert mask_predict_schedules(N, DEFAULT_MP_CONFIG) == tuple(O.mask_predict_schedules(N, O.DEFAULT_MP_CONFIG)) or \
            list(mask_predict_schedules(N, DEFAULT_MP_CONFIG)) == list(O.mask_predict_schedules(N, O.DEFAULT_MP_CONFIG))
    n, temp = mask_predict_schedules(2048, DEFAULT_MP_CONFIG)
    assert len(n) == 50 and len(temp) == 50 and n[0] == int(2048 * 0.9) and n[10] == 256 and n[20] == 128


def _face_reference(grid, vc_mode, face_mode, MASK):
    # independent statement of dalle_bert.py:796-848 on a [b,t,8,8] grid
    out = torch.full_like(grid, MASK)
    if vc_mode == "face_8x8":
        if face_mode == "eyes_nose":
            out[:, :, 2:5, 1:7] = grid[:, :, 2:5, 1:7]
        else:
            out[:, :, 5:7, 2:6] = grid[:, :, 5:7, 2:6]
    elif vc_mode == "face2_8x8":
        out[:, 0] = grid[:, 0]
        out[:, 1:, 2:6, 2:6] = grid[:, 1:, 2:6, 2:6]
    elif vc_mode == "face3_8x8":
        out[:, 0] = grid[:, 0]
        out[:, :, 2:6, 2:6] = grid[:, :, 2:6, 2:6]
    elif vc_mode == "mask_8x8":
        out[:, :, 1:7, 1:7] = grid[:, :, 1:7, 1:7]
    return out


@pytest.mark.parametrize("vc_mode,face_mode", [("face_8x8", "eyes_nose"), ("face_8x8", "mouth"), ("face2_8x8", "x"),
                                               ("face3_8x8", "x"), ("mask_8x8", "x")])
def test_erase_codebook_face_windows(vc_mode, face_mode):
    cfg = dict(BERT_CASES["bert_tiny"], image_size=128, num_visuals=2)
    m, _ = build_bert(cfg, device="cpu")
    ids = torch.randint(0, 1024, (3, 2 * 64))
    got = m.erase_codebook_face(ids.clone(), vc_mode, face_mode).view(3, 2, 8, 8)
    assert torch.equal(got, _face_reference(ids.view(3, 2, 8, 8), vc_mode, face_mode, 1024))


def test_random_erase_half_and_artv_pad_ids():
    cfg = dict(BERT_CASES["bert_tiny"], image_size=128)
    m, _ = build_bert(cfg, device="cpu")
    ids = torch.randint(0, 1024, (2, 64))
    out = m.random_erase_codebook(ids.clone(), None, erase_half=True).view(2, 1, 8, 8)
    assert (out[:, :, 4:] == 1024).all() and torch.equal(out[:, :, :4], ids.view(2, 1, 8, 8)[:, :, :4])
    from cases import ARTV_CASES
    a, _ = build_artv(ARTV_CASES["artv_tiny"], device="cpu")
    assert a._allowed_range(0) == (0, a.num_text_tokens)
    assert a._allowed_range(a.text_seq_len) == (a.num_text_tokens, a.num_control_tokens)
    assert a._allowed_range(a.control_seq_len) == (a.num_control_tokens, a.total_tokens)


def test_training_forward_has_no_cpu_fallback():
    m, _ = build_bert(BERT_CASES["bert_tiny_nov"], device="cpu")
    # the training path runs on the CUDA kernels only: on CPU tensors it must fail loudly, never compute with torch
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 5, dtype=torch.long), target=torch.zeros(2, 3, 3, 32, 32), return_loss=True)


def test_msm_mask_sampler_strategies_and_shapes():
    import random
    m, _ = build_bert(BERT_CASES["bert_tiny"], device="cpu")
    np.random.seed(0); random.seed(0); torch.manual_seed(0)
    for probs in ([1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]):
        mask, nfm = m._sample_msm_masks(3, torch.device("cpu"), np.array(probs, dtype=float), [0.2, 0.5], 0.0)
        assert mask.shape == (3, m.target_seq_len) and mask.dtype == torch.bool
        if probs[1] == 1:  # strategy 2: everything masked, flagged as fully masked
            assert not mask.any() and float(nfm.sum()) == 0
        else:
            assert float(nfm.sum()) == 3
    # pc_prob = 1: at least one whole frame is kept as context
    mask, _ = m._sample_msm_masks(2, torch.device("cpu"), np.array([0, 1.0, 0, 0]), [0.2, 0.5], 1.0)
    per_frame = mask.view(2, m.num_targets, m.image_seq_len).all(dim=2)
    assert per_frame.any(dim=1).all()


def test_shard_bounds_cover_batch_exactly():
    from mmvid_b200.parallel import shard_bounds
    for n in (1, 7, 8, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_stdout_carries_only_the_json_line():
    """bench.py's contract: rank 0 prints ONE JSON line.  Python prints, C-level writes to fd 1 (NCCL's version banner) and
    child processes must end up on stderr once quiet_stdout() ran; emit() writes the line to the saved descriptor."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, os, subprocess\n"
        f"sys.path.insert(0, {root!r})\n"
        "import bench\n"
        "bench.quiet_stdout()\n"
        "print('python noise')\n"
        "os.write(1, b'C-level noise\\n')\n"
        "subprocess.run(['echo', 'child noise'])\n"
        "bench.emit({'metric': 'm', 'value': 1.5})\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert json.loads(r.stdout) == {"metric": "m", "value": 1.5} and r.stdout.count("\n") == 1
    for noise in ("python noise", "C-level noise", "child noise"):
        assert noise in r.stderr


def test_gemm_tile_picker_on_the_benchmark_shapes():
    """Host-only dispatch of mmvid_linear (no GPU needed: 148 SMs are assumed without a device).  2000 + BN = CTA-pair kernel
    with a 256 x BN tile, 1000 + BN = single-CTA kernel.  The benchmark's layers must land on the tiles the pipeline traces
    selected (profiles/r1_g_gemm_pipeline.md); small problems stay on small single-CTA tiles; the 16-bit kinds (bf16 / fp16
    operands, 16-bit or fp32 results through TMA stores) take the widest CTA-pair tile that quantises well over 74 pairs."""
    from mmvid_b200 import _lib as L
    lib = L.load()
    TF32, BF16, F16, F32, B16, H16 = L.TF32, L.BF16, L.F16, L.DT_F32, L.DT_BF16, L.DT_F16
    pick = lib.mmvid_debug_pick_tile
    M = 4 * 2115
    assert pick(M, 3072, 768, TF32, F32) == 2256       # c_fc
    assert pick(M, 768, 3072, TF32, F32) == 2192       # c_proj: 136 tiles of 256 x 192 for 74 pairs instead of 102 of 256 x 256
    assert pick(M, 768, 768, TF32, F32) == 2192        # out_proj
    assert pick(4 * 2048, 1024, 768, TF32, F32) == 2256  # logits head
    for kind, dt16 in ((BF16, B16), (F16, H16)):
        assert pick(M, 3072, 768, kind, dt16) == 2256      # c_fc, 16-bit GELU output
        assert pick(M, 768, 3072, kind, F32) == 2192       # c_proj: 16-bit operands, fp32 residual stream
        assert pick(M, 768, 768, kind, F32) == 2192        # out_proj
        assert pick(4 * 2048, 1024, 768, kind, F32) == 2256  # logits head
    assert pick(130, 192, 128, TF32, F32) == 1064      # tiny problem: most CTAs
    assert pick(0, 192, 128, TF32, F32) // 1000 == 1


def test_fma_pipe_exp2_polynomial_accuracy_claims():
    """Bit-level numpy restatement of exp2_poly2 (tc_attention.cu): clamp at -126, round-to-nearest through the 1.5 * 2^23
    magic constant, minimax polynomial on [-0.5, 0.5], exponent added as an integer.  Pins the accuracy the kernel's header
    claims (degree 4: 2.7e-6 for the tf32 path, degree 3: 7.5e-5 for the bf16 path) and the behaviour at the edges."""
    import numpy as np
    f32 = np.float32
    coef = {4: [0.9999992847442627, 0.6931217908859253, 0.240247443318367, 0.05591785907745361, 0.009570101276040077],
            3: [0.9999280571937561, 0.6932609677314758, 0.2426111251115799, 0.0551716648042202]}

    def exp2_poly(x, deg):
        x = np.maximum(x.astype(f32), f32(-126.0))
        t = (x + f32(12582912.0)).astype(f32)
        f = (x - (t - f32(12582912.0)).astype(f32)).astype(f32)
        p = np.full_like(f, f32(coef[deg][-1]))
        for c in coef[deg][-2::-1]:
            p = (p * f + f32(c)).astype(f32)
        bits = p.view(np.uint32) + (t.view(np.uint32) << np.uint32(23))
        return bits.view(f32)

    x = np.linspace(-60.0, 8.9, 2_000_001).astype(f32)   # scores after max subtraction, up to the lazy-rescale slack
    ref = np.exp2(x.astype(np.float64))
    for deg, bound in ((4, 3.0e-6), (3, 8.0e-5)):
        rel = np.abs(exp2_poly(x, deg).astype(np.float64) / ref - 1.0)
        assert rel.max() < bound, (deg, rel.max())
    # masked keys (-inf) and very negative scores give a tiny positive number instead of an exponent-field borrow
    edge = exp2_poly(np.array([-np.inf, -1e30, -126.0, -125.7], dtype=f32), 4)
    assert np.all(edge >= 0) and np.all(edge < 3e-38)


def test_visual_aug_motion_color_follows_the_reference_rng_order_and_semantics():
    """augment_visual = dalle_bert.py:940-944 + warp_video_with_color (:140-158): python random gates (p = 0.9), then per clip
    torch.rand(1) and random.randint(0, 3); first control frame untouched; other modes are ignored without drawing."""
    import random
    import torch
    from mmvid_b200.augment import augment_visual
    g = torch.Generator().manual_seed(0)
    visual = torch.rand(3, 4, 3, 8, 8, generator=g)
    random.seed(5)
    torch.manual_seed(5)
    out = augment_visual(visual, "motion_color")
    # restatement with the reference's own statements
    random.seed(5)
    torch.manual_seed(5)
    exp = visual
    if random.random() < 0.9:
        exp = visual.detach().clone()
        clips = []
        for n in range(visual.shape[0]):
            x = visual[n, 1:]
            c_shift = torch.rand(1) - 0.5
            m = torch.zeros_like(x)
            num = random.randint(0, 3)
            if num == 0:
                m += c_shift
            else:
                m[:, num - 1] += c_shift
            clips.append(torch.clamp(x + m, 0, 1))
        exp[:, 1:] = torch.stack(clips)
    assert torch.equal(out, exp)
    assert torch.equal(out[:, 0], visual[:, 0]) and not torch.equal(out[:, 1:], visual[:, 1:])
    state = random.getstate()
    assert augment_visual(visual, "something_else") is visual and random.getstate() == state
    assert augment_visual(visual, None) is visual
