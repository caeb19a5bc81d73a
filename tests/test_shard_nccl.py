"""Sample-sharded ingest over a real process group (the north star's collective): different samples on different ranks, one
all-gather of the per-bundle splice signatures, identical bundle_group::resolve on every rank, results compared with the
single-device run (tests/shard_worker.py).  CPU tier: world_size 2 over gloo on the kernel-logic build.  GPU tier: torchrun with
one rank per GPU over NCCL (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_shard_nccl.py -m gpu`)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, emu, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import shard_worker
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = shard_worker.run("gloo", lib_path=emu, samples=4, templates=8000)
        q.put((rank, r, None))
    except Exception as e:      # noqa: BLE001
        q.put((rank, None, repr(e)))
    dist.destroy_process_group()


def test_sample_sharded_resolve_gloo(emu_lib):
    world = 2
    port = _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_gloo_worker, args=(r, world, port, emu_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, r, err in res:
        assert err is None, "rank %d: %s" % (rank, err)
    r0 = [r for rank, r, _ in res if rank == 0][0]
    assert r0["clusters"] > 0 and r0["bytes_received"] > 0


@pytest.mark.gpu
def test_sample_sharded_resolve_nccl(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    world = 2
    out = str(tmp_path / "shard.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "shard_worker.py"), "nccl", out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    r = json.load(open(out))
    assert r["backend"] == "nccl" and r["world"] == world and r["clusters"] > 0 and r["bytes_received"] > 0
