"""One-process-per-GPU driver of the density pass (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

The path shards by POSITIONS of the context's spatial order: every rank builds the same deterministic
order from the replicated coordinates, scans its block-cyclic shard of the positions (blocks of 1024,
dealt round-robin: dense and sparse regions are spread over all ranks) against all frames, and
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
    """rows per shard after padding: whole row blocks, the same on every rank (dcb200_shard_capacity)."""
    blocks = (n + ROW_BLOCK - 1) // ROW_BLOCK
    return (blocks + world_size - 1) // world_size * ROW_BLOCK


def shard_positions(n, world_size, rank):
    """positions of a block-cyclic shard, in the shard's own order: blocks rank, rank + world_size, ... of ROW_BLOCK
    positions each (the dealing of dcb200_ctx_populations_shard for n_cols <= 16; used by the CPU tests)."""
    blocks = range(rank, (n + ROW_BLOCK - 1) // ROW_BLOCK, world_size)
    return [p for b in blocks for p in range(b * ROW_BLOCK, min(n, (b + 1) * ROW_BLOCK))]


def shard_bounds(n, world_size, rank):
    """contiguous shards of shard_size positions (the dealing used on the GEMM-form path, n_cols >= 17)."""
    per = shard_size(n, world_size)
    return min(n, rank * per), min(n, (rank + 1) * per)


def gather_shards(local, world_size, out=None):
    """local: [k][capacity] on every rank -> [world_size][k][capacity] on every rank, ONE all-gather."""
    if world_size == 1:
        return local.unsqueeze(0)
    if out is None:
        out = torch.empty((world_size,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    # concatenation along dim 0 is the layout both NCCL and gloo accept: [world_size * k][capacity]
    dist.all_gather_into_tensor(out.view((world_size * local.shape[0],) + tuple(local.shape[1:])), local)
    return out


def shards_to_position_order(gathered, n, cyclic=True):
    """gathered: [world_size][k][capacity] -> [k][n] in position order; the mapping of the library's
    shards_to_frame_order / nn_finish_shards kernels without the final position -> frame scatter (tests, host logic)."""
    w, k, cap = gathered.shape
    pos = torch.arange(n, device=gathered.device)
    if cyclic:
        blk = pos // ROW_BLOCK
        shard, local = blk % w, (blk // w) * ROW_BLOCK + pos % ROW_BLOCK
    else:
        shard, local = pos // cap, pos % cap
    return gathered[shard, :, local].transpose(0, 1).contiguous()


class _OnSessionStream:
    """Runs a block on the session's CUDA stream, ordered after the caller's current stream on entry and before it on
    exit, so that callers on torch's default stream need no explicit synchronisation (the session stream is non-blocking)."""

    def __init__(self, session):
        self.s = session

    def __enter__(self):
        self.outer = torch.cuda.current_stream(self.s.dev)
        self.inner = self.s.torch_stream()
        self.inner.wait_stream(self.outer)
        self.ctx = torch.cuda.stream(self.inner)
        self.ctx.__enter__()
        return self.inner

    def __exit__(self, *exc):
        self.ctx.__exit__(*exc)
        self.outer.wait_stream(self.inner)
        return False


class DensityPass:
    """populations -> free energies -> nearest neighbours of one trajectory on this rank's GPU, sharded
    over the ranks of the default process group.  `session` is a clustering_b200.session.Session.

    Every rank builds the same deterministic order from the replicated coordinates and scans its block-cyclic shard
    of the positions against all frames; the two data-path collectives are ONE all-gather of the shard populations
    and ONE of the shard neighbour keys.  `timings` (optional dict) receives CUDA events around the stages."""

    def __init__(self, session, n, radii):
        self.s = session
        self.n = n
        self.radii = radii
        self.world, self.rank = world()
        dev = session.dev
        r = len(radii)
        cap = shard_size(n, self.world) if self.world > 1 else n
        self.cap = cap
        self.pops_loc = torch.zeros((r, cap), dtype=torch.int32, device=dev)
        self.keys_loc = torch.zeros((2, cap), dtype=torch.int64, device=dev)
        torch.cuda.current_stream(dev).synchronize()       # the fills must not race with the session's own stream
        if self.world > 1:
            self.pops_all = torch.empty((self.world, r, cap), dtype=torch.int32, device=dev)
            self.keys_all = torch.empty((self.world, 2, cap), dtype=torch.int64, device=dev)
        self.pops_frame = torch.empty((r, n), dtype=torch.int32, device=dev)
        self.fe = torch.empty(n, dtype=torch.float32, device=dev)
        self.events = None

    def _mark(self, name):
        if self.events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.s.dev))
            self.events.append((name, ev))

    def run(self, coords, fe_radius_index=0, timed=False):
        """coords: device tensor [n][d] or host numpy array.  Returns (pops [R][n], fe [n], nn tuple), frame order.
        timed: keep CUDA events between the stages (stage_ms() after a synchronize)."""
        s = self.s
        self.events = [] if timed else None
        with _OnSessionStream(s):
            self._mark("start")
            s.set_coords(coords)
            self._mark("layout")
            if self.world == 1:
                s.populations(self.radii, 0, self.n, out=self.pops_loc)
                self._mark("pops")
                pops = s.to_frame_order(self.pops_loc, out=self.pops_frame)
            else:
                s.populations_shard(self.radii, self.rank, self.world, out=self.pops_loc)
                self._mark("pops")
                gather_shards(self.pops_loc, self.world, out=self.pops_all)
                pops = s.shards_to_frame_order(self.pops_all, len(self.radii), self.world, out=self.pops_frame)
            fe = s.free_energies(pops[fe_radius_index], out=self.fe)
            s.nn_prepare(fe)
            self._mark("gather+fe+ranks")
            if self.world == 1:
                s.nn_scan(0, self.n, out=self.keys_loc)
                self._mark("nn")
                nn = s.nn_finish(self.keys_loc)
            else:
                s.nn_scan_shard(self.rank, self.world, out=self.keys_loc)
                self._mark("nn")
                gather_shards(self.keys_loc, self.world, out=self.keys_all)
                nn = s.nn_finish_shards(self.keys_all, self.world)
            self._mark("gather+finish")
        return pops, fe, nn

    def stage_ms(self):
        """milliseconds per stage of the last run(timed=True); call after the stream is synchronised."""
        ev = self.events or []
        return {name: ev[k - 1][1].elapsed_time(e) for k, (name, e) in enumerate(ev) if k > 0}


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
        with _OnSessionStream(session):
            session.set_coords(sorted_coords, keep_order=True)

    def step(self, m_prev, m_new, max_dist2, comp):
        """comp: int32 device tensor with at least m_new entries; entries >= m_prev are (re)initialised here."""
        s = self.s
        cuts = screen_cuts(m_prev, m_new, self.world)
        with _OnSessionStream(s):               # everything below is ordered on the session's (non-blocking) stream
            comp[m_prev:m_new] = torch.arange(m_prev, m_new, dtype=comp.dtype, device=comp.device)
            s.screening_scan(m_prev, m_new, max_dist2, comp, cuts[self.rank], cuts[self.rank + 1])
            s.screening_flatten(m_new, comp)
            if self.world > 1:
                mine = comp[:m_new].contiguous()
                others = torch.empty((self.world, m_new), dtype=comp.dtype, device=comp.device)
                dist.all_gather_into_tensor(others, mine)
                for g in range(self.world):
                    if g != self.rank:
                        s.screening_merge(m_new, comp, others[g])
                s.screening_flatten(m_new, comp)
        return comp
