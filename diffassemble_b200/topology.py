"""Piece-graph topology producers (the inputs of the hot path; scope rows a14 / N3).

* ``dense_edge_index(n)``: the fully-connected graph with self loops the reference
  builds with ``dense_to_sparse(ones(n, n))`` (``puzzle_dataset.py:279-284``): row-major
  ``(src, dst)`` pairs.
* ``random_regular_edges`` / ``expander_edge_index``: the Exphander d-regular random graph
  of ``puzzle_dataset.py:33-152`` -- a random permutation joined to its ``d // 2`` cyclic
  shifts (plus a perfect matching when ``d`` is odd), symmetrised.  The same
  ``numpy.random.Generator`` calls are made in the same order, so a given seed yields the
  reference's graph.  The optional spectral-gap retry (``eigsh`` on the Laplacian, up to 5
  draws, keep the best lambda_2) is available with ``check_spectral_gap=True``; note that all
  candidates are relabelled circulant graphs with identical spectra, so that retry only ever
  selects among the first five draws by floating-point noise.
* ``batch_graphs``: PyG ``DataLoader`` collation of topologies (node offsets + batch vector).
"""
import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch


def dense_edge_index(num_nodes: int, device=None) -> torch.Tensor:
    idx = torch.arange(num_nodes, device=device)
    src = idx.repeat_interleave(num_nodes)
    dst = idx.repeat(num_nodes)
    return torch.stack([src, dst])


def resolve_degree(num_nodes: int, degree: Union[int, str]) -> int:
    """``"60%"`` -> round(60 * (n - 1) / 100)  (puzzle_dataset.py:46-47, train_script.py:41-46)."""
    if isinstance(degree, str):
        degree = round((int(degree[:-1]) * (num_nodes - 1)) / 100)
    return int(degree)


def random_regular_edges(num_nodes: int, degree: int, rng: Optional[np.random.Generator] = None) -> Tuple[np.ndarray, np.ndarray]:
    if (num_nodes * degree) % 2 != 0:
        raise TypeError("nodes * degree must be even")
    if rng is None:
        rng = np.random.default_rng()
    if degree == 0:
        return np.array([], dtype=np.int64), np.array([], dtype=np.int64)
    perm = rng.permutation(np.arange(num_nodes))
    half = degree // 2
    # neighbour at cyclic distance s in the permuted ring: position p pairs with position p - s
    pos = np.arange(num_nodes)
    shifts = np.arange(1, half + 1)
    a = np.tile(perm, half)
    b = perm[(pos[None, :] - shifts[:, None]) % num_nodes].reshape(-1)
    if degree % 2:
        m = num_nodes // 2
        a = np.concatenate([a, perm[:m]])
        b = np.concatenate([b, perm[m:]])
    return np.concatenate([a, b]), np.concatenate([b, a])


def _lambda2(senders, receivers, num_nodes):
    """lambda_2 of the unnormalised Laplacian, as ``get_eigenvalue`` (puzzle_dataset.py:106-112):
    float32 weights, self loops dropped, duplicate edges summed, ARPACK ``which="SM"``."""
    from scipy.sparse import csr_matrix, identity
    from scipy.sparse.linalg import eigsh

    keep = senders != receivers
    s, r = senders[keep], receivers[keep]
    adj = csr_matrix((np.ones(len(s), dtype=np.float32), (s, r)), shape=(num_nodes, num_nodes))
    deg = np.bincount(s, minlength=num_nodes).astype(np.float32)
    lap = identity(num_nodes, dtype=np.float32, format="csr").multiply(deg[:, None]).tocsr() - adj
    vals = eigsh(lap.tocoo(), k=2, which="SM", return_eigenvectors=False)
    return vals[0] if len(vals) else 0.0


def expander_edge_index(num_nodes: int, degree: Union[int, str], rng: Optional[np.random.Generator] = None,
                        max_num_iters: int = 5, check_spectral_gap: bool = False) -> torch.Tensor:
    """Returns ``edge_index`` ``[2, E]`` (int64), E = n * d, symmetric, no self loops."""
    degree = resolve_degree(num_nodes, degree)
    if rng is None:
        rng = np.random.default_rng()
    if num_nodes <= degree:
        degree = num_nodes - 1
    if num_nodes <= 10:  # complete graph without self loops (puzzle_dataset.py:68-73)
        idx = np.arange(num_nodes)
        s, r = np.repeat(idx, num_nodes), np.tile(idx, num_nodes)
        keep = s != r
        return torch.from_numpy(np.stack([s[keep], r[keep]])).long()
    bound = max(0, degree - 2 * math.sqrt(degree - 1) - 0.1) if degree > 0 else 0
    best, best_val, val, it = None, -1.0, -1.0, 1
    while val < bound and it <= max_num_iters:
        s, r = random_regular_edges(num_nodes, degree, rng)
        if not check_spectral_gap:
            best = (s, r)
            break
        val = _lambda2(s, r, num_nodes)
        if val > best_val:
            best_val, best = val, (s, r)
        it += 1
    return torch.from_numpy(np.stack(best)).long()


def batch_graphs(edge_indices: Sequence[torch.Tensor], num_nodes: Sequence[int]):
    offs, eis, batch = 0, [], []
    for g, (ei, n) in enumerate(zip(edge_indices, num_nodes)):
        eis.append(ei + offs)
        batch.append(torch.full((n,), g, dtype=torch.long, device=ei.device))
        offs += n
    return torch.cat(eis, dim=1).contiguous(), torch.cat(batch)


def expander_permutations(num_nodes: int, seeds: Sequence[int]) -> np.ndarray:
    """One node permutation per graph, drawn with the reference's generator (``random_regular_edges`` above): the only
    random input of an Exphander graph.  int32 ``[len(seeds), num_nodes]``."""
    return np.stack([np.random.default_rng(s).permutation(np.arange(num_nodes)) for s in seeds]).astype(np.int32)


def expander_batch_from_permutations(perms, degree: Union[int, str], device):
    """``edge_index`` / ``batch`` of a batch of equally sized Exphander graphs, written ON the device from the graphs'
    permutations (``perms``: int32 ``[B, n]``, numpy or torch, host -- pinned host memory copies asynchronously -- or
    device).  Stream-ordered on the current stream: this is what a dataloader hands over instead of the ``[2, E]``
    int64 edge list (115 KB instead of 250 MB for 32 x 900-node graphs at 60 %)."""
    import ctypes as C

    from . import _cabi

    lib = _cabi.load_library()
    device = torch.device(device)
    perm_d = (torch.from_numpy(perms) if isinstance(perms, np.ndarray) else perms).to(device=device, dtype=torch.int32,
                                                                                      non_blocking=True).contiguous()
    num_graphs, num_nodes = perm_d.shape
    degree = resolve_degree(num_nodes, degree)
    if num_nodes <= degree:
        degree = num_nodes - 1
    if num_nodes <= 10:
        raise ValueError("graphs of <= 10 nodes are complete graphs in the reference; use expander_edge_index")
    ei = torch.empty((2, num_graphs * num_nodes * degree), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        st = lib.da_expander_edge_index(C.c_void_p(perm_d.data_ptr()), num_nodes, degree, num_graphs, C.c_void_p(ei[0].data_ptr()),
                                        C.c_void_p(ei[1].data_ptr()), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    if st != _cabi.DA_OK:
        raise _cabi.DiffAssembleError(st, "da_expander_edge_index failed")
    perm_d.record_stream(torch.cuda.current_stream(device))
    batch = torch.arange(num_graphs, device=device).repeat_interleave(num_nodes)
    return ei, batch


class ExpanderBatchSpec:
    """Deferred topology of a batch of Exphander graphs: ``GNN_Diffusion.prefetch`` accepts it in place of
    ``edge_index`` and calls :meth:`build` on its staging stream, so the edge list never exists on the host."""

    def __init__(self, perms, degree: Union[int, str]):
        self.perms = torch.from_numpy(perms) if isinstance(perms, np.ndarray) else perms
        self.degree = degree

    @property
    def num_nodes(self):
        return self.perms.shape[0] * self.perms.shape[1]

    def host_bytes(self):
        return self.perms.numel() * self.perms.element_size()

    def build(self, device):
        return expander_batch_from_permutations(self.perms, self.degree, device)


class DenseBatchSpec:
    """Deferred topology of ``num_graphs`` fully connected ``num_nodes``-node graphs (``dense_to_sparse(ones)`` per graph,
    PyG collation), built on the device."""

    def __init__(self, num_nodes: int, num_graphs: int):
        self.n, self.B = num_nodes, num_graphs

    @property
    def num_nodes(self):
        return self.n * self.B

    def host_bytes(self):
        return 0

    def build(self, device):
        one = dense_edge_index(self.n, device=device)
        offs = (torch.arange(self.B, device=device) * self.n)[:, None, None]
        ei = (one[None] + offs).permute(1, 0, 2).reshape(2, -1).contiguous()
        batch = torch.arange(self.B, device=device).repeat_interleave(self.n)
        return ei, batch


def expander_batch_on_device(num_nodes: int, degree: Union[int, str], num_graphs: int, seeds: Sequence[int], device):
    """Scope row N3: batched Exphander ``edge_index`` / ``batch`` built ON the device.

    The only host work is drawing one permutation per graph with the reference's generator
    (``np.random.default_rng(seed).permutation``); the ``[2, B * n * d]`` int64 edge list (250 MB for 32 x 900-node
    graphs at 60 %) is written by a CUDA kernel in the reference's edge order, so it is bit-identical to
    ``batch_graphs([expander_edge_index(n, d, rng=default_rng(seed)) ...])`` without ever existing on the host.
    (The reference's spectral-gap retry only ever re-draws among isospectral graphs, see above, and is skipped.)"""
    if len(seeds) != num_graphs:
        raise ValueError("one seed per graph")
    return expander_batch_from_permutations(expander_permutations(num_nodes, seeds), degree, device)
