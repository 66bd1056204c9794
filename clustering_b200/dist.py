"""One-process-per-GPU driver of the density pass (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

The path shards by POSITIONS of the context's spatial order: every rank builds the same deterministic
order from the replicated coordinates, scans its contiguous range of positions against all frames, and
the per-shard results are assembled with one all-gather per stage (populations, neighbour keys).
The only data-path collectives are those two all-gathers; everything else is replicated.
"""
import torch
import torch.distributed as dist

ROW_BLOCK = 1024          # rows per CTA work item (clustering_b200/csrc/common.cuh: ROWS_PER_CTA)


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_size(n, world_size):
    """positions per rank: whole row blocks, the same on every rank (the last ranks may run empty)."""
    blocks = (n + ROW_BLOCK - 1) // ROW_BLOCK
    return (blocks + world_size - 1) // world_size * ROW_BLOCK


def shard_bounds(n, world_size, rank):
    per = shard_size(n, world_size)
    return min(n, rank * per), min(n, (rank + 1) * per)


def all_gather_positions(local, n, world_size, rank, out=None):
    """local: [k][rows_of_this_rank] (position order) -> [k][n] on every rank.

    Shards are padded to the common shard size so that one all_gather_into_tensor moves everything."""
    k = local.shape[0]
    per = shard_size(n, world_size)
    if world_size == 1:
        return local
    b, e = shard_bounds(n, world_size, rank)
    send = torch.zeros((k, per), dtype=local.dtype, device=local.device)
    send[:, :e - b] = local
    recv = torch.empty((world_size * k, per), dtype=local.dtype, device=local.device)     # ranks concatenated along dim 0
    dist.all_gather_into_tensor(recv, send)
    full = recv.view(world_size, k, per).permute(1, 0, 2).reshape(k, world_size * per)[:, :n]
    if out is not None:
        out.copy_(full)
        return out
    return full.contiguous()


class DensityPass:
    """populations -> free energies -> nearest neighbours of one trajectory on this rank's GPU, sharded
    over the ranks of the default process group.  `session` is a clustering_b200.session.Session."""

    def __init__(self, session, n, radii):
        self.s = session
        self.n = n
        self.radii = radii
        self.world, self.rank = world()
        self.b, self.e = shard_bounds(n, self.world, self.rank)
        dev = session.dev
        r = len(radii)
        self.pops_loc = torch.zeros((r, self.e - self.b), dtype=torch.int32, device=dev)
        self.keys_loc = torch.zeros((2, self.e - self.b), dtype=torch.int64, device=dev)
        self.pops_frame = torch.empty((r, n), dtype=torch.int32, device=dev)
        self.fe = torch.empty(n, dtype=torch.float32, device=dev)

    def run(self, coords, fe_radius_index=0):
        """coords: device tensor [n][d] or host numpy array.  Returns (pops [R][n], fe [n], nn tuple), frame order."""
        s = self.s
        s.set_coords(coords)
        s.populations(self.radii, self.b, self.e, out=self.pops_loc)
        pops_pos = all_gather_positions(self.pops_loc, self.n, self.world, self.rank)
        pops = s.to_frame_order(pops_pos, out=self.pops_frame)
        fe = s.free_energies(pops[fe_radius_index], out=self.fe)
        s.nn_prepare(fe)
        s.nn_scan(self.b, self.e, out=self.keys_loc)
        keys = all_gather_positions(self.keys_loc, self.n, self.world, self.rank)
        nn = s.nn_finish(keys)
        return pops, fe, nn
