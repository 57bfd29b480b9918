"""Multi-GPU sharding of a batch of independent units.

Units (LZNT1 chunks, independent Xpress / Xpress-Huffman buffers) have no cross-unit data dependence, so a batch is cut
into contiguous ranges, one per rank; every rank runs the single-GPU pipeline on its range; the ONLY exchange is one
all-gather of the per-rank output sizes so that each rank can place (or report) its output in the global stream.
Works with any torch.distributed backend: NCCL on the B200s, gloo in the CPU tests.
"""


def partition(count, world, rank):
    """Contiguous unit range [lo, hi) of `rank`: floor(rank*count/world) .. floor((rank+1)*count/world)."""
    return (rank * count) // world, ((rank + 1) * count) // world


def gather_sizes(local_total, device=None, group=None):
    """All-gather of one int64 per rank.  Returns (sizes[world], exclusive offset of this rank, grand total)."""
    import torch
    import torch.distributed as dist
    if not torch.is_tensor(local_total):
        local_total = torch.tensor([int(local_total)], dtype=torch.int64, device=device)
    local_total = local_total.reshape(1).to(torch.int64)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_total.clone(), 0, int(local_total.item())
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=local_total.device)
    dist.all_gather_into_tensor(sizes, local_total, group=group)
    host = sizes.cpu()
    return sizes, int(host[:rank].sum().item()), int(host.sum().item())


def torch_broadcaster(dist, device):
    """`broadcast` argument of GpuContext.comm_init over an initialised torch.distributed group (rank 0 is the source)."""
    import torch

    def broadcast(payload):
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if payload is not None:
            t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())
    return broadcast
