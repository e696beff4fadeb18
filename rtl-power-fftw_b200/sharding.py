"""Multi-GPU decomposition of one integration.

Spectra are independent and the only cross-spectrum operation is a sum
(reference src/datastore.cxx:84), so the stream is sharded by *spectrum index*
-- never by raw buffer, because buffers may straddle spectra
(src/datastore.cxx:52,68) -- and the per-GPU float64 pwr[] vectors are combined
by one allreduce(sum) at the end of the integration.  One process per GPU;
torch.distributed provides the collective (NCCL on GPUs, gloo in CPU tests)."""


def shard_range(total_spectra, rank, world):
    """Contiguous, balanced range [first, last) of spectrum indices for `rank`."""
    if world < 1 or not (0 <= rank < world) or total_spectra < 0:
        raise ValueError("bad shard arguments")
    first = (total_spectra * rank) // world
    last = (total_spectra * (rank + 1)) // world
    return first, last


def shard_byte_range(n_bins, total_spectra, rank, world):
    """Byte range of the u8 IQ stream that `rank` consumes (2 bytes per complex sample)."""
    first, last = shard_range(total_spectra, rank, world)
    return 2 * n_bins * first, 2 * n_bins * last


def allreduce_pwr(pwr_tensor, repeats_done, group=None):
    """Sum pwr[] (float64 tensor, in place) and repeats_done over all ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return pwr_tensor, int(repeats_done)
    dist.all_reduce(pwr_tensor, op=dist.ReduceOp.SUM, group=group)
    r = torch.tensor([int(repeats_done)], dtype=torch.int64, device=pwr_tensor.device)
    dist.all_reduce(r, op=dist.ReduceOp.SUM, group=group)
    return pwr_tensor, int(r.item())
