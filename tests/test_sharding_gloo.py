"""CPU, world_size 2 over gloo: the multi-GPU host logic of config 5 (contiguous sharding of
the global game-id range + one all_reduce(SUM) of the tallies).  The per-rank kernel is
replaced by the C oracle here (this container has no GPU); on the GPU box the same
sharded_sweep runs qttt_sweep under NCCL (bench.py, tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_sweep(lo, hi, seed, device):
    from oracle import c_oracle as CO
    stats, hist = CO.selfplay(lo, hi, seed)
    return torch.from_numpy(np.concatenate([stats, hist]).astype(np.int64))


def _worker(rank, world, port, total, seed, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import qtttgym_b200 as Q
        lo, hi = Q.shard_range(total, rank, world)
        stats = Q.sharded_sweep(total, seed, device="cpu", sweep_fn=_oracle_sweep)
        if rank == 0:
            np.save(out_path, np.array(stats.tolist() + [lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_sweep_world_size_2(tmp_path):
    total, seed = 30001, 77
    out = str(tmp_path / "stats.npy")
    mp.spawn(_worker, args=(2, _free_port(), total, seed, out), nprocs=2, join=True)
    got = np.load(out)
    want = _oracle_sweep(0, total, seed, "cpu").numpy()
    assert got[:16].tolist() == want.tolist()          # identical to the single-process result
    assert got[16:].tolist() == [0, total // 2]         # rank 0 owns the first half


def test_shard_range_partitions_exactly():
    import qtttgym_b200 as Q
    for total in (0, 1, 7, 125_000_000, 10**12 + 3):
        for world in (1, 2, 3, 4, 8):
            cuts = [Q.shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
