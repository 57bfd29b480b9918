"""World-size-2 `gloo` test of the multi-GPU host logic (unit partition + the single all-gather of output sizes)."""
import os
import socket
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_partition_covers_everything(ms):
    from importlib import import_module
    sh = import_module("mscompress_b200.sharding")
    for count in (0, 1, 7, 8, 65536, 262144 + 3):
        for world in (1, 2, 3, 8):
            ranges = [sh.partition(count, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_gather_sizes_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        from importlib import import_module
        import mscompress_b200
        sh = import_module("mscompress_b200.sharding")
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        lo, hi = sh.partition(1000, world, rank)
        local = sum(100 + i for i in range(lo, hi))            # pretend unit i compresses to 100+i bytes
        sizes, off, total = sh.gather_sizes(local)
        assert total == sum(100 + i for i in range(1000)), total
        assert off == sum(100 + i for i in range(0, lo)), off
        assert sizes.tolist()[rank] == local
        dist.destroy_process_group()
        sys.stdout.write("rank " + str(rank) + " ok" + chr(10)); sys.stdout.flush()   # one write: the ranks share a pipe
    """ % ROOT))
    last = None
    for attempt in range(3):                    # (the port is chosen before torchrun binds it: retry if somebody else took it meanwhile)
        port = free_port()
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        last = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                               "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=240)
        if last.returncode == 0:
            break
    assert last.returncode == 0, last.stdout + last.stderr
    assert "rank 0 ok" in last.stdout and "rank 1 ok" in last.stdout
