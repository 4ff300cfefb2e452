"""Column sharding of msa2eds across ranks (one process per GPU): which columns a rank owns and holds, and
the one exchange step of the path — an all-gather of the ranks' output byte counts, from which every rank
gets the offsets of its slices in the single .eds / .seds pair. Plumbing only: the data
path has no collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) moves 16 bytes per rank."""
import os


def plan(total_cols, world, rank, halo):
    """(own_begin, own_end, win_begin, win_end) of `rank`: contiguous equal column ranges plus a halo."""
    lo = total_cols * rank // world
    hi = total_cols * (rank + 1) // world
    return lo, hi, max(0, lo - halo), min(total_cols, hi + halo)


def gather_offsets(dist, device, eds_bytes, seds_bytes, scratch=None):
    """All-gather (eds_bytes, seds_bytes) of every rank; returns (eds_offset, seds_offset, eds_total, seds_total).
    With world size 1 (or an uninitialised process group) nothing is exchanged."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 0, eds_bytes, seds_bytes
    world, rank = dist.get_world_size(), dist.get_rank()
    if scratch is None:
        scratch = (torch.zeros(2, dtype=torch.int64, device=device), torch.zeros(2 * world, dtype=torch.int64, device=device))
    mine, everyone = scratch
    mine[0], mine[1] = int(eds_bytes), int(seds_bytes)
    dist.all_gather_into_tensor(everyone, mine)
    counts = everyone.view(world, 2).cpu()
    eds_off = int(counts[:rank, 0].sum())
    seds_off = int(counts[:rank, 1].sum())
    return eds_off, seds_off, int(counts[:, 0].sum()), int(counts[:, 1].sum())


class OffsetExchange:
    """The same exchange without a host round trip per call: byte counts go to the device through a pinned
    staging tensor, the all-gather is issued asynchronously behind the transform (the next transform on this rank
    does not wait for the other ranks), and the offsets are read back only when the caller asks (`offsets()`),
    e.g. once before the file writes. Two exchanges may be in flight (ring of two buffer sets). Used by bench.py's
    timed loop."""

    RING = 2

    def __init__(self, dist, device):
        import torch

        self.dist = dist
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.host = [torch.zeros(2, dtype=torch.int64) for _ in range(self.RING)]
        if device.type == "cuda":
            self.host = [h.pin_memory() for h in self.host]
        self.mine = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(self.RING)]
        self.everyone = [torch.zeros(2 * self.world, dtype=torch.int64, device=device) for _ in range(self.RING)]
        self.work = [None] * self.RING
        self.posted = 0
        self.last = (0, 0)
        # numpy views of the pinned staging tensors: a plain store instead of a tensor indexing op per value
        self.host_np = [h.numpy() for h in self.host]

    def post(self, eds_bytes, seds_bytes):
        self.last = (int(eds_bytes), int(seds_bytes))
        if self.world == 1:  # nothing to exchange: the offsets are (0, 0) and the totals are this rank's counts
            self.posted += 1
            return
        k = self.posted % self.RING
        if self.work[k] is not None:
            self.work[k].wait()  # the exchange that used this buffer set two posts ago
            self.work[k] = None
        self.host_np[k][0], self.host_np[k][1] = self.last
        self.mine[k].copy_(self.host[k], non_blocking=True)
        self.work[k] = self.dist.all_gather_into_tensor(self.everyone[k], self.mine[k], async_op=True)
        self.posted += 1

    def flush(self):
        """Make the current stream wait for every exchange still in flight (no host synchronisation)."""
        for w in self.work:
            if w is not None:
                w.wait()
        self.work = [None] * self.RING

    def offsets(self):
        if self.world == 1:
            return 0, 0, self.last[0], self.last[1]
        k = (self.posted - 1) % self.RING
        self.flush()
        counts = self.everyone[k].view(self.world, 2).cpu()
        return (int(counts[:self.rank, 0].sum()), int(counts[:self.rank, 1].sum()), int(counts[:, 0].sum()),
                int(counts[:, 1].sum()))


def write_slice(path, offset, data, total, rank):
    """Every rank writes its slice of the one output file at its offset (rank 0 sizes the file first)."""
    flags = os.O_WRONLY | os.O_CREAT
    fd = os.open(path, flags, 0o644)
    try:
        if rank == 0:
            os.ftruncate(fd, total)
        view = memoryview(data)
        done = 0
        while done < len(view):  # Linux caps one write at 0x7ffff000 bytes: loop until the slice is out
            n = os.pwrite(fd, view[done:done + (1 << 30)], offset + done)
            if n <= 0:
                raise OSError("short write to %s at offset %d" % (path, offset + done))
            done += n
    finally:
        os.close(fd)
