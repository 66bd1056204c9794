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


# ------------------------------------------------------------------------------------------------
# screening (free-energy-sorted frames; SURVEY.md 8e: the one stage with an exchange per threshold)
# ------------------------------------------------------------------------------------------------
def screen_cuts(m_prev, m_new, world_size):
    """Row ranges [cut[g], cut[g+1]) of the new sorted positions [m_prev, m_new): every rank gets about the same number of
    PAIRS (row p has p candidate columns below it), like the in-process path of dcb200_screening_step."""
    a, b = float(m_prev) ** 2, float(m_new) ** 2
    cuts = [max(m_prev, min(m_new, int((a + (b - a) * g / world_size) ** 0.5))) for g in range(world_size)] + [m_new]
    for g in range(1, world_size + 1):
        cuts[g] = max(cuts[g], cuts[g - 1])
    return cuts


class ScreeningPass:
    """One free-energy threshold at a time, sharded over the ranks of the default process group.

    Every rank holds the free-energy-sorted coordinates (keep_order context) and the replicated union-find forest
    `comp` (int32 [m], comp[p] <= p, roots = smallest sorted position of a cluster).  A step scans this rank's share of
    the new rows against all lower positions, then ONE all-gather moves the per-rank forests and every rank unions them
    on the device (dcb200_ctx_screening_merge) -- instead of the reference's per-sweep H2D / D2H / host merge
    (density_clustering_cuda.cu:505-571)."""

    def __init__(self, session, sorted_coords):
        self.s = session
        self.world, self.rank = world()
        session.set_coords(sorted_coords, keep_order=True)

    def step(self, m_prev, m_new, max_dist2, comp):
        """comp: int32 device tensor with at least m_new entries; entries >= m_prev are (re)initialised here."""
        s = self.s
        comp[m_prev:m_new] = torch.arange(m_prev, m_new, dtype=comp.dtype, device=comp.device)
        cuts = screen_cuts(m_prev, m_new, self.world)
        s.screening_scan(m_prev, m_new, max_dist2, comp, cuts[self.rank], cuts[self.rank + 1])
        s.screening_flatten(m_new, comp)
        if self.world > 1:
            mine = comp[:m_new].contiguous()
            others = torch.empty((self.world, m_new), dtype=comp.dtype, device=comp.device)
            with torch.cuda.stream(s.torch_stream()):
                dist.all_gather_into_tensor(others, mine)
                for g in range(self.world):
                    if g != self.rank:
                        s.screening_merge(m_new, comp, others[g])
                s.screening_flatten(m_new, comp)
        return comp
