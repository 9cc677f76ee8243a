"""world_size-2 gloo tests of the data-parallel plumbing: sharding the batch over ranks and averaging
the gradients reproduces the single-process full-batch gradient of the path (computed with the CPU oracle), the
timing reduction is a max over ranks, and shards tile the batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import hp_oracle as O

KW = dict(patch_size=4, window_size=16, shift_size=4, shift_strategy="nest_roll", rel_pos_bias="flat", embed_dim=16,
          depths=[2, 2], num_heads=[2, 4], use_cos_attn=True, use_v2_norm_placement=True, dim_in=12 * 16 * 16, f_in=3,
          f_out=4, base_pix=12)
GLOBAL_BATCH = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss_and_grads(x, sd, cfg):
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    y = O.hp_unet_forward(x, sd, cfg)
    y.square().mean().backward()
    keys = sorted(k for k, v in sd.items() if v.grad is not None)
    return keys, [sd[k].grad for k in keys]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from heal_swin_b200 import dist as D

    torch.set_num_threads(2)
    r, l, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    cfg = O.HPConfig(**KW)
    sd = O.synth_state_dict(cfg, seed=3)
    x = torch.randn(GLOBAL_BATCH, KW["f_in"], KW["dim_in"], generator=torch.Generator().manual_seed(11))
    lo, hi = D.shard_range(GLOBAL_BATCH, rank, world)
    keys, grads = _loss_and_grads(x[lo:hi], sd, cfg)
    D.allreduce_mean_(grads)
    slow = D.max_over_ranks(10.0 + rank)
    D.barrier()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), slow=slow, lo=lo, hi=hi, **{k: g.numpy() for k, g in zip(keys, grads)})
    torch.distributed.destroy_process_group()


def test_batch_sharding_plus_gradient_allreduce_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cfg = O.HPConfig(**KW)
    sd = O.synth_state_dict(cfg, seed=3)
    x = torch.randn(GLOBAL_BATCH, KW["f_in"], KW["dim_in"], generator=torch.Generator().manual_seed(11))
    keys, full = _loss_and_grads(x, sd, cfg)
    outs = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    assert [(int(o["lo"]), int(o["hi"])) for o in outs] == [(0, 2), (2, 4)]
    for o in outs:
        assert float(o["slow"]) == 11.0  # max over ranks of (10 + rank)
        for k, g in zip(keys, full):
            # equal shards + mean loss: the rank-mean of the shard gradients is the full-batch gradient
            np.testing.assert_allclose(o[k], g.numpy(), rtol=2e-4, atol=1e-7, err_msg=k)
    for k in keys:  # every rank holds the same reduced gradient
        assert np.array_equal(outs[0][k], outs[1][k])


@pytest.mark.parametrize("gb,world", [(64, 8), (10, 4), (3, 2), (1, 2)])
def test_shard_ranges_tile_the_batch(gb, world):
    from heal_swin_b200.dist import shard_range

    edges = [shard_range(gb, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == gb
    assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
