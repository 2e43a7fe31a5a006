"""Multi-GPU sharding of calc_r_K: one process per GPU, strips of element rows.

The reference parallelises the element loop over contiguous chunks of element ids (one per Julia task) and then
sums the per-task vectors and matrices (FiniteElement.jl:88-147). Here a chunk is a strip of element rows owned by
one GPU. Because unknowns are numbered node-major (Mesh.jl:276-284) a strip touches ONE contiguous range of rows of
r and ONE contiguous range of nzval, and the ranges of neighbouring strips overlap exactly in the interface (the
two node rows a strip boundary shares). The "sum over tasks" therefore reduces to: add the overlap slices of the
two neighbours (send/recv), plus one scalar all-reduce for the residual norm. Nothing else crosses GPUs.

All functions work on torch tensors with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import dataclasses


def strip_rows(num2el, world, rank):
    """Element rows [r0, r1) of strip `rank`.

    Every strip must hold at least TWO element rows: a quadratic strip touches node rows e2 .. e2 + 2, so with a
    single-row strip rank k and rank k + 2 would share a node row and the exchange -- which only talks to the ranks
    k - 1 and k + 1 -- would silently drop that overlap (and an empty strip has no range at all)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    if world > 1 and num2el // world < 2:
        raise ValueError(f"{num2el} element rows cannot be cut into {world} strips of at least 2 rows: "
                         "use fewer ranks (the interface exchange only couples neighbouring strips)")
    return (rank * num2el) // world, ((rank + 1) * num2el) // world


def strip_elements(num1el, num2el, world, rank):
    """1-based inclusive element-id range of strip `rank` (element id = e1 + (e2-1)*num1el, Mesh.jl:582-588)."""
    r0, r1 = strip_rows(num2el, world, rank)
    return r0 * num1el + 1, r1 * num1el


def touched_ranges(mesh, colptr, el_first, el_last):
    """What the elements [el_first, el_last] touch, like maf_range_info: (row_lo, row_hi, slot_lo, slot_hi),
    1-based inclusive. `colptr` is the 1-based CSC column pointer of the pattern (maf_pattern)."""
    nodes = mesh.IX[:, el_first - 1:el_last]
    lo, hi = int(nodes.min()), int(nodes.max())
    eqs = mesh.ID[:, lo - 1:hi]
    eqs = eqs[eqs != 0]
    if eqs.size == 0:
        return (1, 0, 1, 0)
    e0, e1 = int(eqs.min()), int(eqs.max())
    return (e0, e1, int(colptr[e0 - 1]), int(colptr[e1]) - 1)


@dataclasses.dataclass
class Overlap:
    peer: int
    rows: slice        # rows of r shared with the peer
    slots: slice       # entries of nzval shared with the peer


def overlaps(ranges, rank):
    """ranges[k] = (row_lo, row_hi, slot_lo, slot_hi) of rank k, 1-based inclusive (maf_range_info).
    Returns the overlaps of `rank` with its two neighbours as 0-based python slices."""
    out = []
    for nb in (rank - 1, rank + 1):
        if 0 <= nb < len(ranges):
            lo_r, hi_r = max(ranges[rank][0], ranges[nb][0]), min(ranges[rank][1], ranges[nb][1])
            lo_s, hi_s = max(ranges[rank][2], ranges[nb][2]), min(ranges[rank][3], ranges[nb][3])
            nr, ns = max(0, hi_r - lo_r + 1), max(0, hi_s - lo_s + 1)
            out.append(Overlap(nb, slice(lo_r - 1, lo_r - 1 + nr), slice(lo_s - 1, lo_s - 1 + ns)))
    return out


def owned_rows(ranges, rank):
    """Rows of r this rank counts in the residual norm: its range minus what the lower neighbour already counts."""
    lo = ranges[rank][0] - 1 if rank == 0 else max(ranges[rank][0] - 1, ranges[rank - 1][1])
    return slice(lo, ranges[rank][1])


def owned_slots(ranges, rank):
    """Entries of nzval this rank hands to the host: its range minus what the lower neighbour already delivers
    (after the interface exchange both neighbours hold the complete sums of the overlap)."""
    lo = ranges[rank][2] - 1 if rank == 0 else max(ranges[rank][2] - 1, ranges[rank - 1][3])
    return slice(lo, ranges[rank][3])


class InterfaceExchange:
    """Sums the interface rows of r and entries of nzval with the neighbouring strips and all-reduces |r|^2."""

    def __init__(self, dist, ranges, rank, like):
        import torch
        self.dist, self.rank = dist, rank
        self.ovs = overlaps(ranges, rank)
        self.own = owned_rows(ranges, rank)
        self.bufs = []
        for ov in self.ovs:
            n = (ov.rows.stop - ov.rows.start) + (ov.slots.stop - ov.slots.start)
            self.bufs.append((torch.empty(n, dtype=like.dtype, device=like.device),
                              torch.empty(n, dtype=like.dtype, device=like.device)))

    def bytes_per_step(self):
        return sum(2 * s.numel() * s.element_size() for s, _ in self.bufs)

    def __call__(self, r, nzval, rnorm2):
        dist = self.dist
        ops = []
        for ov, (sbuf, rbuf) in zip(self.ovs, self.bufs):
            nr = ov.rows.stop - ov.rows.start
            sbuf[:nr].copy_(r[ov.rows])
            sbuf[nr:].copy_(nzval[ov.slots])
            ops.append(dist.P2POp(dist.isend, sbuf, ov.peer))
            ops.append(dist.P2POp(dist.irecv, rbuf, ov.peer))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for ov, (sbuf, rbuf) in zip(self.ovs, self.bufs):
            nr = ov.rows.stop - ov.rows.start
            r[ov.rows] += rbuf[:nr]
            nzval[ov.slots] += rbuf[nr:]
        rnorm2[0] = (r[self.own] ** 2).sum()
        dist.all_reduce(rnorm2)
        return rnorm2
