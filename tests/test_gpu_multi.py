"""The multi-GPU entry of the C ABI (mscomp_gpu_comm_*, mscomp_gpu_lznt1_compress_sharded): two ranks, one GPU each, the
library's own NCCL communicator.  Needs two CUDA devices (skipped otherwise; the round-end scaling run exercises the same path
through bench.py at N = 2, 4, 8)."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

CHILD = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    from importlib import import_module
    import torch, torch.distributed as dist
    import mscompress_b200 as ms
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sh = import_module("mscompress_b200.sharding")
    ctx = ms.GpuContext(local)
    ctx.comm_init(rank, world, sh.torch_broadcaster(dist, dev))
    assert ctx.comm_size() == world
    # one logical stream of 96 MiB of text, cut at chunk boundaries by mscomp_gpu_shard_range
    import ctypes, numpy as np
    total_chunks = 96 * 256
    first, count = ctypes.c_uint64(), ctypes.c_uint64()
    ms.lib().mscomp_gpu_shard_range(ctypes.c_uint64(total_chunks), rank, world, ctypes.byref(first), ctypes.byref(count))
    host = np.empty(96 << 20, dtype=np.uint8)
    ms.generate_batch_into(host, 1 << 20, 96, 42, "text")
    mine = torch.from_numpy(host[first.value * 4096:(first.value + count.value) * 4096]).to(dev)
    n = mine.numel()
    d_comp = torch.empty(n + 2 * (n // 4096) + 64, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n // 4096 + 1, dtype=torch.int64, device=dev)
    d_res = torch.zeros(ms.RESULT_BYTES, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    ctx.lznt1_compress_sharded(mine, d_comp, d_off, d_res, sizes)
    torch.cuda.synchronize()
    my_bytes = ms.parse_result(d_res)["bytes"]
    got = sizes.cpu().tolist()
    assert got[rank] == my_bytes and ctx.collectives() == 1
    # every rank sees the same sizes; the parts in rank order are the stream the single-GPU call makes of the whole input
    ref = torch.tensor(got, dtype=torch.int64, device=dev)
    dist.broadcast(ref, src=0)
    assert ref.cpu().tolist() == got
    if rank == 0:
        whole = torch.from_numpy(host).to(dev)
        c1 = ms.GpuContext(local)
        d_c, d_o, d_r = c1.lznt1_compress(whole)
        torch.cuda.synchronize()
        assert ms.parse_result(d_r)["bytes"] == sum(got)
        assert torch.equal(d_c[:my_bytes], d_comp[:my_bytes])
        print("multi-gpu ok", got)
    dist.destroy_process_group()
""")


@pytest.mark.gpu
def test_two_ranks_shard_an_lznt1_stream_through_the_c_entry(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    script = tmp_path / "child.py"
    script.write_text(CHILD % ROOT)
    for attempt in range(3):
        port = 29600 + (os.getpid() + 17 * attempt) % 300
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
        if r.returncode == 0 or "address already in use" not in (r.stdout + r.stderr).lower():
            break
    assert r.returncode == 0 and "multi-gpu ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
