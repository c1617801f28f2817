"""GPU, 2 ranks over NCCL (needs >= 2 visible GPUs; skipped on a one-GPU box): the real BERT replica per rank, batch split,
one all-gather - the gathered ids and frames equal a single process's output when every sample has its own seed."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from cases import BERT_CASES
    from helpers import build_bert
    from mmvid_b200 import synth
    from mmvid_b200.parallel import generate_images_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = BERT_CASES["bert_tiny"]
    model, _ = build_bert(cfg, device=f"cuda:{rank}", precision="tf32")
    n = 5  # ragged split: 3 + 2
    text = synth.synth_text(n, cfg["text_seq_len"], cfg["vocab"], 7).cuda()
    visual = synth.synth_frames(n, cfg["num_visuals"], cfg["image_size"], 8).cuda()
    seeds = [100 + i for i in range(n)]
    images, _, seq = generate_images_sharded(model, text, visual, sample_seeds=seeds, mask_predict_steps=4, dynamic=False)
    # single-process reference on this rank's own replica (identical weights: key-based synthetic state dict)
    ref_images, ref_seq = [], []
    for i in range(n):
        torch.manual_seed(seeds[i])
        im, _, sq = model.generate_images(text[i:i + 1], visual=visual[i:i + 1], mask_predict_steps=4, dynamic=False)
        ref_images.append(im)
        ref_seq.append(sq)
    ok_seq = bool(torch.equal(seq, torch.cat(ref_seq, 0)))
    ok_img = bool(torch.equal(images, torch.cat(ref_images, 0)))
    q.put((rank, ok_seq, ok_img, tuple(images.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_sharded_generate_images_equals_single_process_with_per_sample_seeds():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, ok_seq, ok_img, shape in res:
        print(f"rank {rank}: gathered ids equal single-process ids: {ok_seq}; frames equal: {ok_img}; shape {shape}")
        assert ok_seq and ok_img and shape[0] == 5
