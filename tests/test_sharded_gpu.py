"""Multi-GPU parity of the sharded reductions (NCCL over NVLink), one process per GPU via torch.distributed.run.
Skipped on boxes with fewer than 2 GPUs (the single-GPU driver run); `gpurun --gpus 2` exercises it."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("no_p2p", ["0", "1"])
def test_sharded_reductions_two_ranks(no_p2p):
    """no_p2p = 0: tiny partials go through the peer-memory mailbox kernel; 1: everything through NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, HPTB_NO_P2P=no_p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + os.getpid() % 500), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-4000:] + "\n" + r.stderr[-4000:]
    assert r.stdout.count("sharded ok") == 2, r.stdout[-2000:]
    if no_p2p == "1":
        assert "peer memory 0" in r.stdout


@pytest.mark.gpu
def test_single_rank_comm_offsets_arg_indices():
    """world = 1: no exchange, but argmax over the shard axis still reports GLOBAL indices (offset added)."""
    import numpy as np
    import hpt_b200 as hb
    ctx = hb.context(0)
    try:
        comm = hb.Comm(ctx, 1, 0, hb.Comm.unique_id())
    except hb.HptError as ex:
        pytest.skip(f"NCCL unavailable: {ex}")
    x = np.random.default_rng(3).standard_normal((9, 5)).astype(np.float32)
    X = hb.ShardedTensor(hb.Tensor.to_cuda(torch.from_numpy(x)), comm, 0, 100, 40)
    got = X.argmax(0).to_cpu().numpy()
    np.testing.assert_array_equal(got, x.argmax(axis=0) + 40)
    np.testing.assert_allclose(X.sum([0]).to_cpu().numpy(), x.sum(axis=0), rtol=1e-5, atol=1e-5)
    comm.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_exchange_protocol_with_virtual_ranks_on_one_gpu(world):
    """The sharded reductions on ONE GPU: `world` virtual ranks (hptb_comm_init_local_group — mailboxes in the same device
    memory, no NCCL, no IPC), one stream per rank, every rank's kernel of a call in flight at once.  Exercises exactly the
    code the multi-GPU runs execute — the exchange fused into reduce_rows_kernel / reduce_cols_lean_kernel, the
    standalone xchg_combine_kernel, index offsets, f32 exchange of half types — at the single-GPU bars, on the box the
    driver has.  (Real NVLink traffic: tests/sharded_worker.py on ≥ 2 GPUs.)
    Runs in a subprocess with CUDA_MODULE_LOADING=EAGER and a timeout: kernels of different streams wait for each other
    here, and with lazy loading the FIRST launch of a kernel may synchronise the context — which a spinning peer kernel
    would block for ever.  (One process per GPU, the real deployment, has no such coupling.)"""
    env = dict(os.environ, CUDA_MODULE_LOADING="EAGER")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "virtual_ranks_worker.py"), str(world)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "virtual ranks ok" in r.stdout, r.stdout[-2000:]
