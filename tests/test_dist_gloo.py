"""CPU: the N>1 sampling path (batch split + single all-gather) on world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeModel:
    """Stands in for BERT on CPU: deterministic 'frames' derived from the prompt ids (no kernels)."""
    num_targets, image_seq_len = 2, 4

    def generate_images(self, text, visual=None, **kw):
        b = text.shape[0]
        frames = text.float().sum(1).view(b, 1, 1, 1, 1).expand(b, self.num_targets, 3, 4, 4).contiguous()
        seq = text[:, :1].repeat_interleave(self.num_targets, 0).repeat(1, self.image_seq_len)
        return frames, [], seq


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmvid_b200.parallel import generate_images_sharded
    text = torch.arange(n * 5).view(n, 5)
    images, _, seq = generate_images_sharded(_FakeModel(), text)
    ref_images, _, ref_seq = _FakeModel().generate_images(text)
    q.put((rank, bool(torch.equal(images, ref_images)), bool(torch.equal(seq, ref_seq)), tuple(images.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5])
def test_sharded_generation_gathers_global_batch(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    for rank, ok_i, ok_s, shape in res:
        assert ok_i and ok_s and shape[0] == n


class _SeededFake(_FakeModel):
    """'Sampling' that consumes the global torch RNG: only per-sample seeding makes its output partition-independent."""

    def generate_images(self, text, visual=None, **kw):
        b = text.shape[0]
        frames = torch.rand(b, self.num_targets, 3, 4, 4) + text.float().sum(1).view(b, 1, 1, 1, 1)
        seq = torch.randint(0, 1024, (b * self.num_targets, self.image_seq_len))
        return frames, [], seq


def _seeded_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmvid_b200.parallel import generate_images_sharded
    text = torch.arange(n * 5).view(n, 5)
    seeds = [50 + 3 * i for i in range(n)]
    images, _, seq = generate_images_sharded(_SeededFake(), text, sample_seeds=seeds)
    ref_i, ref_s = [], []
    for i in range(n):
        torch.manual_seed(seeds[i])
        im, _, sq = _SeededFake().generate_images(text[i:i + 1])
        ref_i.append(im)
        ref_s.append(sq)
    q.put((rank, bool(torch.equal(images, torch.cat(ref_i))), bool(torch.equal(seq, torch.cat(ref_s)))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5])
def test_per_sample_seeds_make_sharded_sampling_partition_independent(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_seeded_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(a and b for _, a, b in res)


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmvid_b200.parallel import all_reduce_gradients
    g = torch.Generator().manual_seed(3)
    shapes = [(5, 7), (3,), (11, 2, 2), (1,)]
    base = [torch.randn(*s, generator=g) for s in shapes]
    params = [torch.nn.Parameter(torch.zeros(*s)) for s in shapes]
    for i, (p, b) in enumerate(zip(params, base)):
        if not (rank == 1 and i == 2):            # rank 1 has no gradient for parameter 2
            p.grad = b * (rank + 1)
    all_reduce_gradients(params, bucket_bytes=64)  # tiny buckets: several collectives
    ok = True
    for i, (p, b) in enumerate(zip(params, base)):
        want = b * (1.5 if i != 2 else 0.5)         # mean of (1, 2) x b; parameter 2: (1 + 0) / 2
        ok = ok and bool(torch.allclose(p.grad, want, atol=1e-6))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gradient_all_reduce_averages_over_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(ok for _, ok in res)
