"""Oracle restatement of the graph-topology producers (test infrastructure).

Follows ``puzzle_diff/dataset/puzzle_dataset.py:33-152`` (random d-regular
expander by cyclic shifts of a random permutation, with the spectral-gap retry)
and ``:279-284`` (dense graph = ``dense_to_sparse(ones(N, N))``: row-major
``(src, dst)`` pairs including self loops), plus PyG ``DataLoader`` batching
(per-graph node offsets on ``edge_index``, a ``batch`` vector).
"""
import math

import numpy as np
import torch


def dense_edge_index(num_nodes: int) -> torch.Tensor:
    """``pyg.utils.dense_to_sparse(torch.ones(N, N))[0]`` == ``ones.nonzero().t()``."""
    return torch.ones(num_nodes, num_nodes).nonzero().t().contiguous()


def generate_random_regular_graph(num_nodes, degree, rng=None):
    # puzzle_dataset.py:115-152
    if (num_nodes * degree) % 2 != 0:
        raise TypeError("nodes * degree must be even")
    if rng is None:
        rng = np.random.default_rng()
    if degree == 0:
        return np.array([]), np.array([])
    nodes = rng.permutation(np.arange(num_nodes))
    num_reps = degree // 2
    num_nodes = len(nodes)
    ns = np.hstack([np.roll(nodes, i + 1) for i in range(num_reps)])
    edge_index = np.vstack((np.tile(nodes, num_reps), ns))
    if degree % 2 != 0:
        edge_index = np.hstack((edge_index, np.vstack((nodes[: num_nodes // 2], nodes[num_nodes // 2 :]))))
    senders = np.concatenate([edge_index[0], edge_index[1]])
    receivers = np.concatenate([edge_index[1], edge_index[0]])
    return senders, receivers


def get_eigenvalue(senders, receivers, num_nodes):
    # puzzle_dataset.py:106-112: unnormalised Laplacian L = D - A (self loops removed,
    # duplicate edges summed), two smallest-magnitude eigenvalues via ARPACK.
    from scipy.sparse import coo_matrix
    from scipy.sparse.linalg import eigsh

    s = np.asarray(senders, dtype=np.int64)
    r = np.asarray(receivers, dtype=np.int64)
    keep = s != r
    s, r = s[keep], r[keep]
    w = np.ones(len(s), dtype=np.float32)
    deg = np.zeros(num_nodes, dtype=np.float32)
    np.add.at(deg, s, w)
    rows = np.concatenate([s, np.arange(num_nodes)])
    cols = np.concatenate([r, np.arange(num_nodes)])
    vals = np.concatenate([-w, deg])
    L = coo_matrix((vals, (rows, cols)), shape=(num_nodes, num_nodes))
    return eigsh(L, k=2, which="SM", return_eigenvectors=False)


def generate_random_expander(num_nodes, degree, rng=None, max_num_iters=5, check_spectral_gap=True):
    # puzzle_dataset.py:33-103
    if isinstance(degree, str):
        degree = round((int(degree[:-1]) * (num_nodes - 1)) / 100)
    if rng is None:
        rng = np.random.default_rng()
    eig_val = -1
    eig_val_lower_bound = max(0, degree - 2 * math.sqrt(degree - 1) - 0.1) if degree > 0 else 0
    max_eig_val_so_far = -1
    max_senders, max_receivers = [], []
    cur_iter = 1
    if num_nodes <= degree:
        degree = num_nodes - 1
    if num_nodes <= 10:
        for i in range(num_nodes):
            for j in range(num_nodes):
                if i != j:
                    max_senders.append(i)
                    max_receivers.append(j)
    else:
        while eig_val < eig_val_lower_bound and cur_iter <= max_num_iters:
            senders, receivers = generate_random_regular_graph(num_nodes, degree, rng)
            if not check_spectral_gap:
                max_senders, max_receivers = senders, receivers
                break
            eig_val = get_eigenvalue(senders, receivers, num_nodes=num_nodes)
            eig_val = 0 if len(eig_val) == 0 else eig_val[0]
            if eig_val > max_eig_val_so_far:
                max_eig_val_so_far = eig_val
                max_senders, max_receivers = senders, receivers
            cur_iter += 1
    max_senders = torch.as_tensor(np.asarray(max_senders), dtype=torch.long).view(-1, 1)
    max_receivers = torch.as_tensor(np.asarray(max_receivers), dtype=torch.long).view(-1, 1)
    return torch.cat([max_senders, max_receivers], dim=1)  # [E, 2]; callers use .t()


def batch_graphs(edge_indices, num_nodes):
    """PyG ``Batch.from_data_list`` for topology only: offset each graph's
    ``edge_index`` by the running node count and build the ``batch`` vector."""
    offs, eis, batch = 0, [], []
    for g, (ei, n) in enumerate(zip(edge_indices, num_nodes)):
        eis.append(ei + offs)
        batch.append(torch.full((n,), g, dtype=torch.long))
        offs += n
    return torch.cat(eis, dim=1).contiguous(), torch.cat(batch)
