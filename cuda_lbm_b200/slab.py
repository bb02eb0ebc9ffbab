"""y-slab decomposition of the D2Q9 domain over the GPUs of one box.

One process per GPU (torchrun); slab g owns rows [y0, y0+ny_local).  The in-place AA pattern touches
the neighbour slab only on odd steps: before such a step a rank needs, per face, the three
populations of the neighbour's edge row that stream into it (3*nx floats), and after it the three
populations it wrote for the neighbour travel back.  torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests) moves those rows; everything else is local.  The other collectives
are the 3-double all-reduce of CM<2,OptimalAdapter>'s grid sums and, when immersed bodies are present,
the all-reduce of the stencil-node states (5 floats per node touched by a marker; lbm_ibm_pack/unpack).

The engine is passed in as an object with the halo/step methods of cuda_lbm_b200.solver.Engine so that
the exchange schedule can be tested on CPU with a stand-in (tests/test_slab_gloo.py).
"""
import torch
import torch.distributed as dist


def neighbours(rank, world, periodic_y):
    """(lower-y rank, upper-y rank) or None where the slab touches a non-periodic domain edge."""
    if world == 1:
        return None, None
    lo = rank - 1 if rank > 0 else (world - 1 if periodic_y else None)
    hi = rank + 1 if rank < world - 1 else (0 if periodic_y else None)
    return lo, hi


def slab_rows(ny, rank, world):
    """Row range of a slab — same rule as lbm_create (engine.cu): the first ny % world slabs get one extra row."""
    base, rem = divmod(ny, world)
    n = base + (1 if rank < rem else 0)
    y0 = rank * base + min(rank, rem)
    return y0, n


class SlabExchange:
    """Halo exchange schedule for one rank.  `engine.halo(what, side, ptr)` packs / unpacks 3*nx floats."""

    def __init__(self, engine, nx, periodic_y, device, rank=None, world=None, group=None):
        self.e = engine
        self.group = group
        if rank is None or world is None:
            on = dist.is_available() and dist.is_initialized()
            rank = dist.get_rank(group) if on else 0
            world = dist.get_world_size(group) if on else 1
        self.rank, self.world = rank, world
        self.lo, self.hi = neighbours(self.rank, self.world, periodic_y)
        self.send = [torch.zeros(3 * nx, dtype=torch.float32, device=device) for _ in range(2)]
        self.recv = [torch.zeros(3 * nx, dtype=torch.float32, device=device) for _ in range(2)]
        self.bytes_per_exchange = sum(3 * nx * 4 for p in (self.lo, self.hi) if p is not None)

    def exchange(self, phase):
        """phase 'pre': edge rows -> neighbours' ghost rows; 'post': ghost rows -> neighbours' edge rows."""
        peers = (self.lo, self.hi)
        for side in (0, 1):
            if peers[side] is not None:
                self.e.halo("pack_" + phase, side, self.send[side].data_ptr())
        ops = []
        # a message packed on my side s is consumed on the peer's side 1-s; when both faces touch the same peer
        # (world == 2, periodic) messages match in posting order, hence receives are posted in reverse side order
        for side in (0, 1):
            if peers[side] is not None:
                ops.append(dist.P2POp(dist.isend, self.send[side], peers[side], group=self.group))
        for side in (1, 0):
            if peers[side] is not None:
                ops.append(dist.P2POp(dist.irecv, self.recv[side], peers[side], group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for side in (0, 1):
            if peers[side] is not None:
                self.e.halo("unpack_" + phase, side, self.recv[side].data_ptr())


def attach_peers(engine, periodic_y, group=None):
    """Peer-mapped coupling (include/lbm_b200.h, lbm_peer_*): every rank exports the descriptor of its population
    buffer, all-gathers them and maps ALL slabs (lbm_peer_attach_all; the two y-neighbours among them are attached as such).
    Afterwards the fused kernel reads / writes the neighbour's edge rows over NVLink itself, CM<2,OptimalAdapter>'s grid sums
    are all-reduced on the device, bodies may span any number of slabs, and `engine.step(n)` needs no host-side exchange."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    descs = [None] * world
    dist.all_gather_object(descs, engine.peer_export(), group=group)
    engine.peer_attach_all(descs)
    dist.barrier(group=group)
    return neighbours(rank, world, periodic_y)


class SlabSolver:
    """The solver loop of one rank.  mode "direct": neighbours' edge rows are peer-mapped (NVLink loads/stores inside the
    fused kernel, device-side step handshake, device-side all-reduce of the adapter sums) and n steps are enqueued at once;
    mode "nccl": lbm_step one step at a time with packed halo rows (and IBM node states, adapter sums) sent through
    torch.distributed in between."""

    def __init__(self, engine, nx, periodic_y, device, optimal_adapter=False, adapter_exact=True, group=None, mode="nccl"):
        self.e = engine
        self.x = SlabExchange(engine, nx, periodic_y, device, group=group)
        self.world = self.x.world
        self.optimal, self.exact = optimal_adapter, adapter_exact
        self.group = group
        self._sums = torch.zeros(3, dtype=torch.float64, device=device)
        self._device = device
        self._ibm = None            # exchange buffer of the IBM node states (sized per step: bodies may be added or moved)
        self.collectives = 0
        self.mode = mode if self.world > 1 else "single"
        if self.mode == "direct":
            attach_peers(engine, periodic_y, group)

    def barrier_after_init(self):
        """All slabs must hold their initial state before any of them steps (the handshake counters restart at 0)."""
        self.e.sync()
        if self.world > 1:
            dist.barrier(group=self.group)

    def _allreduce_sums(self):
        s = torch.tensor(self.e.moment_sums(), dtype=torch.float64, device=self._sums.device)
        dist.all_reduce(s, group=self.group)
        self.e.set_moment_sums(s.cpu().tolist())
        self.collectives += 1

    def _exchange_ibm(self):
        """Bodies across slab faces, halo coupling: every slab contributes the states of the stencil nodes it owns (zeros
        elsewhere), so the sum over slabs is a gather — bit-exact, whatever order the all-reduce adds in."""
        nf = getattr(self.e, "ibm_exchange_floats", lambda: 0)()      # changes when bodies are added or moved
        if self._ibm is None or self._ibm.numel() != nf:
            self._ibm = torch.zeros(nf, dtype=torch.float32, device=self._device)
        if nf == 0:
            return
        self.e.ibm_pack(self._ibm.data_ptr())
        dist.all_reduce(self._ibm, group=self.group)
        self.e.ibm_unpack(self._ibm.data_ptr())
        self.collectives += 1

    def run_from_host(self, rho_ptr, u_ptr, n, rho_out_ptr, u_out_ptr):
        """One driver segment per rank: host rho,u of this slab -> n steps -> host rho,u (integer addresses, ideally pinned memory).
        One slab, or peer-mapped slabs without OptimalAdapter: lbm_run_from_host (copies hidden behind the kernels, the slab faces
        synchronised level by level on the device).  Otherwise the three calls with the barrier the handshake needs."""
        if self.world == 1 or (self.mode == "direct" and not self.optimal):
            self.e.run_from_host(rho_ptr, u_ptr, n, rho_out_ptr, u_out_ptr)
            return
        # OptimalAdapter needs grid sums (no band pipeline), or the halo coupling: the three calls
        self.e.init_fields_local(rho_ptr, u_ptr)
        self.barrier_after_init()
        self.step(n, macroscopics=True)
        self.e.macroscopics_into(rho_out_ptr, u_out_ptr)

    def step(self, n=1, macroscopics=False):
        if self.world == 1 or self.mode == "direct":
            # peer-mapped slabs: halo rows, IBM node states and the adapter's grid sums all travel inside the kernels
            self.e.step(n, macroscopics=macroscopics)
            return
        for i in range(n):
            need = self.e.next_step_needs_halo()
            if need:
                self.x.exchange("pre")
            self._exchange_ibm()
            # exact mode: the grid sums of this step's post-stream state; lagged mode: only while no previous step has produced
            # them (first step after init / restart / set_populations)
            if self.optimal and (self.exact or getattr(self.e, "adapter_sums_pending", lambda: False)()):
                self.e.adapter_prepass()
                self._allreduce_sums()
            self.e.step(1, macroscopics=macroscopics and i == n - 1)
            if self.optimal and not self.exact:
                self._allreduce_sums()
            if need:
                self.x.exchange("post")
