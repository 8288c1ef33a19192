"""Multi-GPU plumbing: one process per GPU, ciphertext batches sharded by contiguous ranges, keys
replicated once, no collective in the steady-state loop.  Only torch.distributed plumbing lives here."""
import torch
import torch.distributed as dist


def shard_range(count, rank, world):
    """Contiguous [lo, hi) slice of `count` independent units owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(count), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_bytes(tensor_u8, src=0):
    """One-time key replication: broadcast a raw byte blob from `src` (NCCL over NVLink on GPUs, gloo on CPU)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor_u8, src=src)
    return tensor_u8


def device_blob_as_tensor(ptr, nbytes, device):
    """Wrap an engine-owned device allocation as a uint8 torch tensor without copying (for dist.broadcast)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


def replicate_gate_keys(engine, params, bk_host=None, ks_host=None, device=None):
    """Rank 0 ingests host keys (coefficient domain) and transforms them on its GPU; every other rank allocates and
    receives the two device blobs (bk spectra, repacked ks) by broadcast.  Single process: plain load."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        engine.load_gate_keys(params, bk_host, ks_host)
        return
    if rank == 0:
        engine.load_gate_keys(params, bk_host, ks_host)
    else:
        engine.alloc_gate_keys(params)
    for which in (0, 1):
        ptr, nbytes = engine.gate_key_blob(which)
        broadcast_bytes(device_blob_as_tensor(ptr, nbytes, device), src=0)
    torch.cuda.synchronize()
    if rank != 0:
        engine.commit_gate_keys()          # only now do the receivers' buffers hold key material


def replicate_cb_keys(engine, params, preKS_host=None, bk_host=None, privKS_host=None, device=None, with_privks=True):
    """Circuit-bootstrap keys, same protocol: rank 0 ingests (device-side transform / repack), the three device blobs -- bk spectra
    131 MB, preKS 37 MB, privKS 2.35 GB at the reference's parameters -- go to the other ranks by broadcast over NVLink."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1 or rank == 0:
        engine.load_cb_keys(params, preKS_host, bk_host, privKS_host if with_privks else None)
        if world == 1:
            return
    else:
        engine.alloc_cb_keys(params, with_privks)
    for which in (0, 1, 2) if with_privks else (0, 1):
        ptr, nbytes = engine.cb_key_blob(which)
        broadcast_bytes(device_blob_as_tensor(ptr, nbytes, device), src=0)
    torch.cuda.synchronize()
    if rank != 0:
        engine.commit_cb_keys()


def max_over_ranks(value, device="cpu"):
    """Timing rule: a multi-GPU number is the MAX over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
