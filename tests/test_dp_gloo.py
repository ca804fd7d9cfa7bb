"""World-size-2 CPU tests (gloo) of the host-side multi-GPU logic:

* image sharding of the extraction path (rank = image_index % world, no collective);
* the data-parallel training protocol -- what each rank puts in the flat buffer, what is summed,
  and how the Adam step normalises it -- which the GPU path (mc_mlp_partial_fit + NCCL) follows
  step for step.  The single-process oracle it is compared with is pinned to the reference
  (tests/test_oracle_head.py).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mermaid_classifier_b200.sharding import images_for_rank, merge_rank_outputs
from mermaid_classifier_b200.torch_classifier import split_steps
from oracle import head as ohead


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cluster(n, d, k, seed):
    rng = np.random.RandomState(seed)
    centers = rng.randn(k, d) * 3.0
    y = rng.randint(0, k, size=n)
    return (centers[y] + rng.randn(n, d) * 1.3).astype(np.float32), y


def _train_rank(rank, world, port, X, y, cw, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    w, b = ohead.init_mlp(X.shape[1], (16, 8), 5, 0)
    adam = ohead.AdamState(w + b)
    curve = [ohead.partial_fit(w, b, adam, X, y, lr=1e-3, random_state=0, class_weight=cw, rank=rank, world=world,
                               all_reduce=lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) for _ in range(2)]
    # every rank must hold identical parameters after identical all-reduced updates
    flat = torch.cat([p.reshape(-1) for p in w + b])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        out.put((curve, [p.numpy() for p in w + b], same, adam.t))
    dist.destroy_process_group()


@pytest.mark.parametrize("weighted", [False, True])
def test_data_parallel_protocol_matches_single_process(weighted):
    X, y = _cluster(650, 32, 5, 42)  # 3 full mini-batches + ragged 50 (25 rows per rank in the tail)
    cw = torch.tensor([0.5 + 0.5 * i for i in range(5)]) if weighted else None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_rank, args=(r, 2, port, X, y, cw, q)) for r in range(2)]
    for p in procs:
        p.start()
    curve, params, same, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w, b = ohead.init_mlp(32, (16, 8), 5, 0)
    adam = ohead.AdamState(w + b)
    want = [ohead.partial_fit(w, b, adam, X, y, lr=1e-3, random_state=0, class_weight=cw) for _ in range(2)]
    assert same and t == adam.t == 8
    np.testing.assert_allclose(curve, want, rtol=1e-5)
    for got, ref in zip(params, w + b):
        np.testing.assert_allclose(got, ref.numpy(), atol=1e-5, rtol=1e-4)


def test_split_steps_partitions_every_minibatch():
    for n, mb, world in [(650, 200, 2), (1000, 200, 8), (7, 200, 4), (401, 200, 3), (0, 200, 2)]:
        seen = []
        per_step = None
        for r in range(world):
            pos, off = split_steps(n, mb, r, world)
            assert off[0] == 0 and off[-1] == len(pos) and np.all(np.diff(off) >= 0)
            assert len(off) - 1 == (n + mb - 1) // mb   # every rank takes part in every step (all-reduce!)
            sizes = np.diff(off)
            per_step = sizes if per_step is None else per_step + sizes
            seen.append(pos)
        allpos = np.sort(np.concatenate(seen)) if seen else np.zeros(0)
        assert np.array_equal(allpos, np.arange(n))
        if n:
            want = [min(mb, n - s) for s in range(0, n, mb)]
            assert per_step.tolist() == want


def _shard_rank(rank, world, port, n_images, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = images_for_rank(n_images, rank, world)
    # stand-in for the per-image feature rows: row value = image index (no collective on the data path)
    feats = {i: np.full((3, 4), i, dtype=np.float32) for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, feats))   # host-side bookkeeping only (counters / manifest)
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_image_sharding_covers_all_images_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_rank, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [m for m, _ in gathered] == [[0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9]]
    merged = merge_rank_outputs([f for _, f in gathered], 11)
    assert [int(a[0, 0]) for a in merged] == list(range(11))
    for world in (1, 2, 4, 8):
        cover = sorted(i for r in range(world) for i in images_for_rank(100, r, world))
        assert cover == list(range(100))


def test_row_sharding_is_a_balanced_contiguous_partition():
    """C4 scoring shards feature rows in contiguous blocks (bench.py's c4_scoring record): the blocks tile [0, n) in rank
    order, sizes differ by at most one row, degenerate sizes (fewer rows than ranks, zero rows) included, and concatenating
    per-rank labels in rank order reproduces the single-process order."""
    from mermaid_classifier_b200.sharding import rows_for_rank

    for n in (0, 1, 7, 8, 9, 10_000_000, 10_000_003):
        for world in (1, 2, 4, 8):
            blocks = [rows_for_rank(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
    labels = np.arange(1003) % 17
    parts = [labels[slice(*rows_for_rank(len(labels), r, 4))] for r in range(4)]
    assert np.array_equal(np.concatenate(parts), labels)
    with pytest.raises(ValueError):
        rows_for_rank(10, 2, 2)


# ---- trainer-level reductions of the throughput (sharded) mode ---------------------------------------------------
class _Group:
    """What trainer.MermaidTrainer reads from torch_classifier.DataParallel (which itself needs NCCL + a GPU)."""

    def __init__(self, rank, world):
        self.rank, self.world, self.group, self.device = rank, world, None, "cpu"


def _trainer_rank(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mermaid_classifier_b200.trainer import MermaidTrainer

    tr = MermaidTrainer(batch_size=100, data_parallel=_Group(rank, world), dp_mode="throughput")
    assert tr._sharded
    # evaluation counts: rank r saw (10 + r) hits, loss sum 1.5 * (r + 1), 100 + r rows
    hits, loss, n = tr._sum_over_ranks(10 + rank, 1.5 * (rank + 1), 100 + rank, "cpu")
    # ragged row gather in rank order (what calibration feeds to the Platt fit)
    rows = torch.arange((3 + 2 * rank) * 4, dtype=torch.float64).reshape(3 + 2 * rank, 4) + 100.0 * rank
    gathered = tr._gather_rows(rows)
    y = tr._gather_rows(torch.full((3 + 2 * rank,), rank, dtype=torch.int32))
    # parity mode and single-process trainers leave everything alone
    plain = MermaidTrainer(batch_size=100, data_parallel=_Group(rank, world), dp_mode="parity")
    untouched = plain._sum_over_ranks(7, 0.25, 9, "cpu") == (7, 0.25, 9) and plain._gather_rows(rows) is rows
    if rank == 0:
        out.put((hits, loss, n, gathered.numpy(), y.numpy(), untouched))
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_sharded_reductions_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    hits, loss, n, gathered, y, untouched = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (hits, loss, n) == (21, 4.5, 201) and untouched
    want = np.concatenate([np.arange(12, dtype=np.float64).reshape(3, 4), np.arange(20, dtype=np.float64).reshape(5, 4) + 100.0])
    assert np.array_equal(gathered, want)
    assert y.tolist() == [0, 0, 0, 1, 1, 1, 1, 1]


class _FakeClf:
    classes_ = np.arange(3)

    def __init__(self):
        self.fits = 0

    def partial_fit(self, x, y, classes=None):
        self.fits += 1


class _HostLabels:
    """ImageLabels stand-in: `n_chunks` host chunks per epoch."""

    def __init__(self, n_chunks):
        self.n_chunks, self.label_count = n_chunks, 100 * n_chunks

    def load_data_in_batches(self, batch_size, random_seed=None):
        for i in range(self.n_chunks):
            yield [[float(i)] * 4] * 5, [0, 1, 2, 0, 1]


def _chunk_rank(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mermaid_classifier_b200.trainer import MermaidTrainer

    tr = MermaidTrainer(batch_size=100, data_parallel=_Group(rank, world), dp_mode="throughput")
    # equal chunk counts: every rank fits all of its chunks
    clf = _FakeClf()
    tr._train_epoch(clf, _HostLabels(3), [0, 1, 2], epoch=0)
    equal_ok = clf.fits == 3
    # unequal counts (rank 1 draws one chunk more): BOTH ranks raise after the common prefix instead of hanging
    clf2 = _FakeClf()
    try:
        tr._train_epoch(clf2, _HostLabels(2 + rank), [0, 1, 2], epoch=0)
        raised = False
    except ValueError:
        raised = True
    # final validation lists are concatenated in rank order on every rank
    g, e = tr._gather_lists([f"gt{rank}a", f"gt{rank}b"], [rank, rank + 10])
    out.put((rank, equal_ok, raised, clf2.fits, g, e))
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_chunk_agreement_and_list_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chunk_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, equal_ok, raised, fits, g, e in res:
        assert equal_ok and raised and fits == 2
        assert g == ["gt0a", "gt0b", "gt1a", "gt1b"] and e == [0, 10, 1, 11]


def test_trainer_rejects_unknown_dp_mode():
    from mermaid_classifier_b200.trainer import MermaidTrainer

    with pytest.raises(ValueError, match="dp_mode"):
        MermaidTrainer(batch_size=10, dp_mode="fast")
