"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank takes its contiguous shard of one batch,
produces that shard's results (here with the CPU oracle standing in for the device, since the product
path has no CPU fallback), and the shards must tile the single-process result exactly; the step time
is the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from fips204_b200.shard import all_shards, shard_range


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 7, 8, 1000, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            sh = all_shards(n, world)
            assert sh[0][0] == 0 and sh[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [hi - lo for lo, hi in sh]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from synth import synth_block
    from fips204_b200.shard import max_over_ranks, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    xi = synth_block("gloo-xi", n, 32)[lo:hi]
    pk, sk = O.keygen_batch(44, xi, nthreads=1)
    msgs = [bytes(x) for x in synth_block("gloo-msg", n, 32)[lo:hi]]
    sig, _ = O.sign_batch(44, sk, msgs, synth_block("gloo-rnd", n, 32)[lo:hi], nthreads=1)
    np.save(os.path.join(out_dir, f"sig{rank}.npy"), sig)
    # no data-path collective: only the barrier and the max-over-ranks time cross ranks
    dist.barrier()
    t = max_over_ranks(10.0 + rank, dist)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt)
    assert t == 10.0 + world - 1 and int(cnt.item()) == n
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process(tmp_path):
    import oracle_lib as O
    from synth import synth_block
    n, world = 13, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"sig{r}.npy") for r in range(world)])
    pk, sk = O.keygen_batch(44, synth_block("gloo-xi", n, 32), nthreads=1)
    want, _ = O.sign_batch(44, sk, [bytes(x) for x in synth_block("gloo-msg", n, 32)], synth_block("gloo-rnd", n, 32),
                           nthreads=1)
    assert got.shape == want.shape and (got == want).all()
