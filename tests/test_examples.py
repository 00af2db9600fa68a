"""The training example runs on gloo, learns nothing exotic, and its checkpoint / resume path is exact."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

from dist_utils import run_distributed  # noqa: E402


def _train_worker(rank, world, argv, out_path):
    import train_ring_transformer as ex

    loss = ex.train(ex.parse_args(argv))
    if rank == 0 and out_path:
        torch.save(torch.tensor(loss), out_path)


COMMON = ["--device", "cpu", "--seq-len", "65", "--dim", "32", "--depth", "2", "--heads", "4", "--kv-heads", "2",
          "--dim-head", "8", "--vocab", "32", "--batch", "2", "--lr", "1e-2", "--log-every", "100"]


@pytest.mark.parametrize("world,replicas", [(2, 1), (4, 2)])
def test_training_example_checkpoint_resume_is_exact(tmp_path, world, replicas):
    extra = ["--batches-per-ring", str(replicas)]
    straight, resumed, ckpt = tmp_path / "a.pt", tmp_path / "b.pt", tmp_path / "ckpt.pt"
    run_distributed(_train_worker, world, COMMON + extra + ["--steps", "4"], str(straight))
    run_distributed(_train_worker, world, COMMON + extra + ["--steps", "2", "--ckpt", str(ckpt), "--ckpt-every", "2"], "")
    assert ckpt.exists() and torch.load(ckpt)["step"] == 2
    run_distributed(_train_worker, world, COMMON + extra + ["--steps", "4", "--ckpt", str(ckpt), "--ckpt-every", "2"],
                    str(resumed))
    a, b = torch.load(straight).item(), torch.load(resumed).item()
    assert a == a and abs(a - b) < 1e-5, (a, b)


def test_training_example_loss_goes_down(tmp_path):
    out = tmp_path / "loss.pt"
    run_distributed(_train_worker, 2, COMMON + ["--steps", "60", "--task", "count"], str(out), timeout=400.0)
    import math

    assert torch.load(out).item() < 0.6 * math.log(32)  # 5 % jumps: the floor is ~0.37 nats


def _decode_worker(rank, world, argv, out_path):
    import decode_tree_attention as ex

    worst = ex.run(ex.parse_args(argv))
    if rank == 0:
        torch.save(torch.tensor(worst), out_path)


@pytest.mark.parametrize("world,context", [(2, 301), (4, 2)])
def test_decode_example_matches_dense_attention(tmp_path, world, context):
    """Sharded-cache decode loop incl. round-robin appends and (context=2, world=4) ranks that start empty."""
    out = tmp_path / "err.pt"
    run_distributed(_decode_worker, world, ["--device", "cpu", "--context", str(context), "--batch", "2", "--heads", "4",
                                            "--kv-heads", "2", "--dim-head", "16", "--steps", "9", "--check"], str(out))
    assert torch.load(out).item() < 1e-4
