"""world_size-2 gloo test of the task-sharded meta-gradient exchange (CPU): the all-reduced packed
buffer equals the serial sum over all tasks, and both replicas end up identical."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from weatherforecast_stgcn_maml_b200.dist import init_from_env, shard_tasks, allreduce_meta
    rank, local, world = init_from_env("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    P, T = 1000, 5
    g = torch.Generator().manual_seed(0)
    per_task = torch.randn(T, P + 4, generator=g, dtype=torch.float64)   # every rank can rebuild the serial sum
    mine = shard_tasks(T, rank, world)
    buf = per_task[mine].sum(0)
    allreduce_meta(buf)
    assert torch.allclose(buf, per_task.sum(0), rtol=0, atol=1e-12), rank
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    assert torch.equal(gathered[0], gathered[1])
    dist.barrier()
    sys.stdout.write("rank " + str(rank) + " ok\\n")   # one write per rank: the two ranks share the pipe
    sys.stdout.flush()
""") % ROOT


def test_gloo_world2_meta_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count(" ok") == 2 and "0" in out.stdout and "1" in out.stdout, out.stdout
