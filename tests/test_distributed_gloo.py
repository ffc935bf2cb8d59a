"""CPU, world_size 2 over gloo: the N>1 exchange step (all_gather of best-hit records +
deterministic merge) gives every rank the single-process answer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparksmithwaterman_b200 import multigpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, scores, cells, lengths, expect):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = multigpu.shard_refs(lengths, rank, world)
    n_reads = scores.shape[1]
    local = np.zeros((n_reads, 4), np.int32)
    for q in range(n_reads):
        k = max(range(len(ids)), key=lambda x: (scores[ids[x], q], -x))
        local[q] = (scores[ids[k], q], k, *cells[ids[k], q])
    merged = multigpu.allgather_best_hits(torch.from_numpy(multigpu.localize(local, ids)))
    assert (merged.numpy() == expect).all(), rank
    dist.destroy_process_group()


def test_allgather_merge_world2():
    rng = np.random.default_rng(5)
    n_refs, n_reads = 40, 17
    scores = rng.integers(0, 30, size=(n_refs, n_reads))
    cells = rng.integers(1, 200, size=(n_refs, n_reads, 2))
    lengths = rng.integers(50, 3000, size=n_refs)
    expect = np.zeros((n_reads, 4), np.int32)
    for q in range(n_reads):
        k = max(range(n_refs), key=lambda x: (scores[x, q], -x))
        expect[q] = (scores[k, q], k, *cells[k, q])
    mp.spawn(_worker, args=(2, _free_port(), scores, cells, lengths, expect), nprocs=2, join=True)


def _worker_empty(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # one reference, two ranks: rank 1's shard is empty and reports (0, -1, 0, 0) for every read
    ids = multigpu.shard_refs([100], rank, world)
    local = np.array([[7, 0, 3, 9], [0, 0, 0, 0]], np.int32) if ids else np.array([[0, -1, 0, 0], [0, -1, 0, 0]], np.int32)
    merged = multigpu.allgather_best_hits(torch.from_numpy(multigpu.localize(local, ids)))
    assert merged.numpy().tolist() == [[7, 0, 3, 9], [0, 0, 0, 0]], (rank, merged)      # never ref -1 for a score-0 read
    dist.destroy_process_group()


def test_empty_shard_never_wins_world2():
    assert multigpu.merge_best_hits(np.array([[[0, -1, 0, 0]], [[0, 0, 0, 0]]])).tolist() == [[0, 0, 0, 0]]
    mp.spawn(_worker_empty, args=(2, _free_port()), nprocs=2, join=True)
