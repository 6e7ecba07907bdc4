"""One IMEX step spread over the GPUs of a box -- the decompositions the problem itself offers:

    world 2, "subdomain": rank 0 owns electrons + holes (semiconductor), rank 1 reductants + oxidants (electrolyte)
                          -- the reference's own split: separate triangulations / DoFHandlers / CarrierPairs
                          (reference include/SolarCell.hpp:351-367)
    world 4, "species"  : one carrier per rank -- the reference's four solve tasks
                          (reference source/SolarCell.cpp:1763-1781)

Per step every rank assembles the right-hand sides of the subdomains it owns a carrier of and solves its carriers
(pecs_step_local), the owners broadcast their new DENSITY blocks (4 of the 12 unknowns per cell; the currents stay
with their owner), and every rank forms the Poisson right-hand side and solves the Poisson system itself
(pecs_step_finish): the potential is needed everywhere and a redundant solve is cheaper than a second exchange.
That broadcast is the only data-path collective: NCCL over NVLink on the context's own stream, so the step stays
asynchronous.  Strong scaling: the work of ONE step is divided.  GpuEngine.connect_p2p() removes even that
collective: the backward sweeps then store every finished density straight into the peers' vectors (CUDA IPC peer
memory over NVLink) while they run, ordered by single-thread flag kernels -- compute and exchange are one kernel.

The driver only needs an ENGINE with step_local(), step_finish(), density(s) -> 1-D torch tensor and
store_density(s, tensor); GpuEngine wraps a SolarCellProblem (tensors alias the context's device memory),
the CPU tests plug in an engine built on the oracle and run the same driver over gloo.
"""
import numpy as np

N_SPECIES = 4


def owner_of(species, world_size):
    """rank that factorises and solves carrier `species`"""
    if world_size not in (1, 2, 4):
        raise ValueError("a step shards over 1, 2 (subdomains) or 4 (species) ranks")
    return species * world_size // N_SPECIES


def owned_mask(rank, world_size):
    return sum(1 << s for s in range(N_SPECIES) if owner_of(s, world_size) == rank)


def mode_name(world_size):
    return {1: "single context", 2: "subdomain: 2 carriers per rank", 4: "species: 1 carrier per rank"}[world_size]


class _DeviceArray:
    """__cuda_array_interface__ view of a block of the context's device memory"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class GpuEngine:
    def __init__(self, prob, device):
        import torch
        self.prob, self.torch = prob, torch
        self.blocks = [torch.as_tensor(_DeviceArray(*prob.density_block(s)), device=f"cuda:{device}") for s in range(N_SPECIES)]
        self.stream = torch.cuda.ExternalStream(prob.stream, device=f"cuda:{device}")

    def step_local(self):
        self.prob.step_local()

    def step_finish(self):
        self.prob.step_finish()

    def density(self, s):
        return self.blocks[s]

    def store_density(self, s, t):
        pass  # the tensor IS the context's memory

    def exchange_context(self):
        return self.torch.cuda.stream(self.stream)  # collectives are ordered on the context's stream

    def connect_p2p(self, dist, rank, world):
        """replace the NCCL broadcasts by stores into the peers' memory fused into the backward sweeps
        (pecs_p2p_export / pecs_p2p_connect); one all_gather_object of the IPC handles at setup"""
        blobs = [None] * world
        dist.all_gather_object(blobs, self.prob.p2p_export())
        self.prob.p2p_connect(rank, world, blobs)
        self.fused_exchange = True


class ShardedStepper:
    def __init__(self, engine, dist, rank, world_size):
        self.engine, self.dist, self.rank, self.world = engine, dist, rank, world_size
        self.owner = [owner_of(s, world_size) for s in range(N_SPECIES)]

    def exchange(self):
        if self.world == 1 or getattr(self.engine, "fused_exchange", False):
            return  # nothing to do between the two halves: the solves have already written into the peers' memory
        ctx = self.engine.exchange_context() if hasattr(self.engine, "exchange_context") else _Null()
        with ctx:
            for s in range(N_SPECIES):
                t = self.engine.density(s)
                self.dist.broadcast(t, src=self.owner[s])
                self.engine.store_density(s, t)

    def step(self, n_steps=1):
        for _ in range(n_steps):
            self.engine.step_local()
            self.exchange()
            self.engine.step_finish()


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def gather_states(prob, dist, rank, world_size, device=None):
    """full state on every rank: each carrier's solution vector from its owner (output / checkpoint path)"""
    import torch
    out = []
    for s in range(N_SPECIES):
        v = prob.get_solution(s)
        if world_size > 1:
            t = torch.from_numpy(np.ascontiguousarray(v))
            if device is not None:
                t = t.to(f"cuda:{device}")
            dist.broadcast(t, src=owner_of(s, world_size))
            v = t.cpu().numpy()
        out.append(v)
    out.append(prob.get_solution(4))
    return out
