"""world_size-2 gloo test of the N>1 host logic: image sharding + the integer confusion-matrix
all-reduce give exactly the single-process result (SURVEY.md §8(e) caveat iv)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ee_semantic_segmentation_b200 import parallel
    from oracle import restate as R
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(99)
    n_img, C = 7, 5
    pred = torch.randint(0, C, (n_img, 64), generator=g)
    tgt = torch.randint(0, C + 1, (n_img, 64), generator=g)
    exits = torch.randint(0, 3, (n_img,), generator=g)
    cm = torch.zeros(4, C + 1, C, dtype=torch.int64)
    counts = torch.zeros(4, dtype=torch.int64)
    for k in parallel.shard_range(n_img, rank, world):
        m = torch.tensor(R.confusion_matrix(pred[k:k + 1].numpy(), tgt[k:k + 1].numpy(), C)[0])
        cm[exits[k]] += m; cm[-1] += m
        counts[exits[k]] += 1; counts[-1] += 1
    parallel.all_reduce_counts(cm, counts)
    torch.save({"cm": cm, "counts": counts, "miou": parallel.miou_from_cm(cm)}, os.path.join(tmp, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_sharded_confusion_allreduce_matches_single_process(tmp_path):
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from ee_semantic_segmentation_b200 import parallel
    from oracle import restate as R
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a["cm"], b["cm"]) and torch.equal(a["counts"], b["counts"])
    g = torch.Generator().manual_seed(99)
    n_img, C = 7, 5
    pred = torch.randint(0, C, (n_img, 64), generator=g)
    tgt = torch.randint(0, C + 1, (n_img, 64), generator=g)
    exits = torch.randint(0, 3, (n_img,), generator=g)
    cm = np.zeros((4, C + 1, C), np.int64)
    for k in range(n_img):
        m = R.confusion_matrix(pred[k:k + 1].numpy(), tgt[k:k + 1].numpy(), C)[0]
        cm[exits[k]] += m; cm[-1] += m
    np.testing.assert_array_equal(a["cm"].numpy(), cm)
    assert int(a["counts"][-1]) == n_img
    tp, fp, fn = R.basics_from_cm(cm[-1])
    ref = (tp / (tp + fp + fn)).sum() / C
    assert float(a["miou"][-1]) == pytest.approx(ref, abs=1e-12)
    assert list(parallel.shard_range(7, 1, 2)) == [1, 3, 5]


def _train_worker(rank, world, port, tmp, buckets=1):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ee_semantic_segmentation_b200 import parallel
    parallel.init_from_env("gloo")
    torch.manual_seed(0)                                   # identical initial weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    fg = parallel.FlatGradients(net.parameters(), buckets=buckets)
    assert len(fg.buckets) == buckets
    g = torch.Generator().manual_seed(7)
    X, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    Xs, ys = X[rank::world], y[rank::world]                # this rank's shard
    for _ in range(3):
        fg.zero()
        if buckets > 1:
            fg.begin()                                     # per-bucket exchange from the gradient hooks, during backward
        ((net(Xs) - ys) ** 2).mean().backward()            # accumulates in place into the flat buffer
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in net.parameters())
        if buckets > 1:
            assert all(v == -1 for v in fg._left)          # every bucket was reduced by its last gradient's hook
            fg.finish()
        else:
            fg.all_reduce_mean()
        opt.step()
    torch.save([p.detach().clone() for p in net.parameters()], os.path.join(tmp, f"p{rank}.pt"))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("buckets", [1, 2])
def test_flat_gradient_allreduce_training_matches_single_process(tmp_path, buckets):
    """The data-parallel step of train_funcs.GraphedTrainStep on CPU/gloo: gradients as views of one flat buffer, one
    all-reduce(mean) per step. Two ranks on disjoint equal shards end with identical parameters, equal to single-process
    SGD on the whole batch (the mean of the shard means is the batch mean)."""
    port = 29500 + ((os.getpid() + 137) % 500)
    mp.spawn(_train_worker, args=(2, port, str(tmp_path), buckets), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "p0.pt"), torch.load(tmp_path / "p1.pt")
    assert all(torch.equal(x, z) for x, z in zip(a, b))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator().manual_seed(7)
    X, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    for _ in range(3):
        opt.zero_grad()
        ((net(X) - y) ** 2).mean().backward()
        opt.step()
    for x, p in zip(a, net.parameters()):
        assert torch.allclose(x, p.detach(), rtol=1e-5, atol=1e-6)
