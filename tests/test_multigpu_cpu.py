"""The N > 1 path on CPU: two gloo ranks shard a request, draw the GLOBAL noise and keep their rows, and all-gather
their results into request order.  Compute is replaced by a deterministic stand-in — the kernels
need a GPU — so this covers exactly the host-side sharding / RNG / collective logic that bench.py uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_predict(batch, lat_shape):
    """stand-in for predict(): consumes the RNG exactly like the real path (posterior c, posterior uc, init noise)"""
    from udifftext_b200.host import rng
    n_c = rng.randn(lat_shape, "cpu")
    n_uc = rng.randn(lat_shape, "cpu")
    x = rng.randn(lat_shape, "cpu")
    img = batch["image"] * 0.25 + (n_c + 2 * n_uc + 3 * x).mean(dim=(1, 2, 3), keepdim=True)
    return img.contiguous()


def _worker(rank, world, port, gb, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from udifftext_b200 import api, synth
    from udifftext_b200.host import rng
    full = synth.synthetic_batch(9, gb, 16, 16, None)
    lo, hi = api.shard_bounds(gb, rank, world)
    mine = api.shard_batch(full, lo, hi)
    assert len(mine["label"]) == hi - lo and mine["image"].shape[0] == hi - lo
    torch.manual_seed(42)
    with rng.batch_shard(gb, lo, hi):
        img = _fake_predict(mine, (hi - lo, 4, 2, 2))
    gathered = api.all_gather_images(img, gb)
    # what bench.py gathers: the uint8 HWC images (udt_images_to_u8 on the GPU; the same expression here), ragged shards too
    u8 = (img.clamp(0, 1).permute(0, 2, 3, 1) * 255).to(torch.uint8).contiguous()
    g8 = api.all_gather_images(u8, gb)
    assert g8.dtype == torch.uint8 and tuple(g8.shape) == (gb, 16, 16, 3)
    assert torch.equal(g8, (gathered.clamp(0, 1).permute(0, 2, 3, 1) * 255).to(torch.uint8))
    torch.save(gathered, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, gb, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, gb, str(tmp_path)), nprocs=world, join=True)
    return [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]


def test_two_ranks_reproduce_the_single_rank_result(tmp_path):
    from udifftext_b200 import synth
    gb = 4
    outs = _run(2, gb, tmp_path)
    full = synth.synthetic_batch(9, gb, 16, 16, None)
    torch.manual_seed(42)
    ref = _fake_predict(full, (gb, 4, 2, 2))
    for o in outs:
        assert o.shape == ref.shape and torch.equal(o, ref)


def test_ragged_shards(tmp_path):
    from udifftext_b200 import synth
    gb = 5
    outs = _run(2, gb, tmp_path)
    full = synth.synthetic_batch(9, gb, 16, 16, None)
    torch.manual_seed(42)
    ref = _fake_predict(full, (gb, 4, 2, 2))
    for o in outs:
        assert torch.equal(o, ref)
