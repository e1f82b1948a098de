"""world_size-2 gloo tests (CPU) of the N>1 host logic: read-balanced sharding + the single all-gather
of per-site outputs reproduce the unsharded arrays exactly."""
import os
import socket
import subprocess
import sys
import textwrap


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from m6anet_b200.dist import all_gather_rows, all_gather_site_outputs, env_world, shard_bounds
    rank, world, _ = env_world()
    dist.init_process_group("gloo")
    rng = np.random.default_rng(7)
    n_reads = rng.integers(20, 400, size=1001)
    site_prob = rng.random(1001).astype(np.float32)
    mod_count = rng.integers(0, 400, size=1001).astype(np.int32)
    bounds = shard_bounds(n_reads, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    sp, mc = all_gather_site_outputs(torch.from_numpy(site_prob[lo:hi].copy()), torch.from_numpy(mod_count[lo:hi].copy()), bounds)
    assert sp.dtype == torch.float32 and mc.dtype == torch.int32
    assert np.array_equal(sp.numpy(), site_prob), "site_prob mismatch"
    assert np.array_equal(mc.numpy(), mod_count), "mod_count mismatch"
    # empty shard on one rank
    b2 = [0, 1001, 1001] if world == 2 else bounds
    lo, hi = b2[rank], b2[rank + 1]
    sp, mc = all_gather_site_outputs(torch.from_numpy(site_prob[lo:hi].copy()), torch.from_numpy(mod_count[lo:hi].copy()), b2)
    assert np.array_equal(sp.numpy(), site_prob) and np.array_equal(mc.numpy(), mod_count)
    # per-pass predictions of validate(): [n_local, n_iterations] rows -> the full matrix on every rank
    y_pred = rng.random((1001, 5)).astype(np.float32)
    for b in (bounds, b2):
        lo, hi = b[rank], b[rank + 1]
        full = all_gather_rows(torch.from_numpy(y_pred[lo:hi].copy()), b)
        assert full.dtype == torch.float32 and np.array_equal(full.numpy(), y_pred), "y_pred mismatch"
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_all_gather_of_site_outputs(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {rank} ok" in out
