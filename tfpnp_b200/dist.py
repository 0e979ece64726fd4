"""Multi-GPU: env_batch shards on dim 0, one process per GPU, state resident per rank.

Replaces ``DataParallelWithCallback(solver)`` (tfpnp/policy/sync_batchnorm/replicate.py:50-75;
tasks/csmri/main.py:79-80), which scatters the state, re-broadcasts all 11.8 M UNet parameters and
gathers the result on EVERY solver call.  The images are independent, so the data path needs no
collective; the only exchange is one all-gather of the per-image PSNR vector (tfpnp/env/base.py:
237-242) after the last iteration.  Works with any torch.distributed backend (NCCL on GPUs, gloo
in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous split of ``n`` items: the first ``n % world`` ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(obj, rank: int, world: int):
    """Slice dim 0 of every tensor in a (nested) dict / list / tuple."""
    if isinstance(obj, torch.Tensor):
        lo, hi = shard_bounds(obj.shape[0], rank, world)
        return obj[lo:hi]
    if isinstance(obj, dict):
        return {k: shard_batch(v, rank, world) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(shard_batch(v, rank, world) for v in obj)
    return obj


def all_gather_batch(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Gather dim-0 shards of any tensor (e.g. ``solver.get_output`` images [B_local,1,H,W]) from all ranks into
    [n_total, ...] in batch order -- the optional second exchange of SURVEY 8e, for callers that need every image on
    every rank (the reference's DataParallel gather, replicate.py:50-75)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def all_gather_psnr(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Gather the [B_local,1] PSNR vectors of all ranks into [n_total,1] (rank order = batch order): the one exchange of the
    data path."""
    return all_gather_batch(local, n_total)


class NativeComm:
    """The NCCL communicator behind the C ABI (``tfpnp_comm_*``, include/tfpnp_b200.h): the PSNR all-gather for
    hosts that do not want a torch.distributed process group on the data path.  The 128-byte ncclUniqueId is created on rank 0
    and handed to the other ranks by the caller (``from_torch_distributed`` uses one broadcast of the existing group for that
    bootstrap; any other channel -- a file, MPI, an environment variable -- works as well)."""

    ID_BYTES = 128

    def __init__(self, unique_id: bytes, rank: int, world: int, device: torch.device):
        import ctypes as C
        from . import _lib
        if len(unique_id) != self.ID_BYTES:
            raise ValueError(f"ncclUniqueId is {self.ID_BYTES} bytes, got {len(unique_id)}")
        self.rank, self.world, self.device = rank, world, device
        self._h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, self.ID_BYTES)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().tfpnp_comm_init(buf, self.ID_BYTES, rank, world, C.byref(self._h)), "tfpnp_comm_init")

    @staticmethod
    def create_unique_id() -> bytes:
        import ctypes as C
        from . import _lib
        buf = C.create_string_buffer(NativeComm.ID_BYTES)
        _lib.check(_lib.lib().tfpnp_comm_unique_id(buf, NativeComm.ID_BYTES), "tfpnp_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: torch.device):
        rank, world = dist.get_rank(), dist.get_world_size()
        on_dev = dist.get_backend() == "nccl"
        t = torch.zeros(cls.ID_BYTES, dtype=torch.uint8, device=device if on_dev else "cpu")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.create_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(bytes(t.cpu().numpy().tobytes()), rank, world, device)

    def all_gather_psnr(self, local: torch.Tensor) -> torch.Tensor:
        """[B_local,1] fp32 on this rank's device -> [world * B_local, 1] (rank order = batch order; equal shards)."""
        from . import _lib
        local = local.contiguous().float()
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=torch.float32, device=local.device)
        with torch.cuda.device(local.device):
            _lib.check(_lib.lib().tfpnp_comm_allgather_psnr(self._h, local.data_ptr(), local.numel(), out.data_ptr(),
                                                            torch.cuda.current_stream().cuda_stream), "tfpnp_comm_allgather_psnr")
        return out

    def __del__(self):
        try:
            from . import _lib
            if self._h:
                _lib.lib().tfpnp_comm_destroy(self._h)
        except Exception:
            pass
