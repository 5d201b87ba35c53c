"""Post-sampling assignment metric (scope row N2), mirroring ``spatial_diffusion.py:179-216, 791-846``.

``greedy_cost_assignment(pos1, pos2)`` has the reference's signature and return value (``[n, 3]`` int64 rows
``(i, j, int64(min_val))`` in greedy order) but runs as one CUDA kernel without host round trips;
``greedy_cost_assignment_batched`` does every puzzle of a batch in one launch (one CTA per puzzle).
"""
import ctypes as C

import torch

from . import _cabi
from ._cabi import DiffAssembleError


def greedy_cost_assignment_batched(pos1: torch.Tensor, pos2: torch.Tensor, graph_ptr: torch.Tensor) -> torch.Tensor:
    """``pos1``, ``pos2``: CUDA fp32 ``[N, >=2]`` (only columns 0, 1 are read; row-strided views are fine);
    ``graph_ptr``: ``[B + 1]`` node offsets.  Returns int64 ``[N, 3]`` (indices local to each puzzle)."""
    lib = _cabi.load_library()
    if not pos1.is_cuda or not pos2.is_cuda:
        raise RuntimeError("greedy_cost_assignment needs CUDA tensors: there is no CPU path")
    pos1, pos2 = pos1.float(), pos2.float()
    if pos1.stride(-1) != 1:
        pos1 = pos1.contiguous()
    if pos2.stride(-1) != 1:
        pos2 = pos2.contiguous()
    gp = graph_ptr.to(device=pos1.device, dtype=torch.int32).contiguous()
    sizes = (graph_ptr[1:] - graph_ptr[:-1])
    max_n = int(sizes.max()) if len(sizes) else 0
    n_total = pos1.shape[0]
    out = torch.empty((n_total, 3), dtype=torch.int64, device=pos1.device)
    if n_total == 0:
        return out
    with torch.cuda.device(pos1.device):
        st = lib.da_greedy_cost_assignment(
            C.c_void_p(pos1.data_ptr()), pos1.stride(0), C.c_void_p(pos2.data_ptr()), pos2.stride(0), C.c_void_p(gp.data_ptr()),
            len(gp) - 1, max_n, C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(pos1.device).cuda_stream))
    if st != _cabi.DA_OK:
        raise DiffAssembleError(st, "da_greedy_cost_assignment failed (puzzle too large for one CTA?)")
    return out


def greedy_cost_assignment(pos1: torch.Tensor, pos2: torch.Tensor) -> torch.Tensor:
    n = pos1.shape[0]
    return greedy_cost_assignment_batched(pos1, pos2, torch.tensor([0, n], dtype=torch.int32))


def real_grid(n_rows: int, n_cols: int, device) -> torch.Tensor:
    """The target cell centres of an ``n_rows x n_cols`` puzzle exactly as ``spatial_diffusion.py:791-794`` builds them
    (``linspace`` per axis, ``meshgrid(x, y, indexing="xy")``, rows flattened first)."""
    y = torch.linspace(-1, 1, n_rows, device=device)
    x = torch.linspace(-1, 1, n_cols, device=device)
    xy = torch.stack(torch.meshgrid(x, y, indexing="xy"), -1)
    return xy.reshape(-1, 2)


@torch.no_grad()
def puzzle_accuracy(img: torch.Tensor, x_gt: torch.Tensor, batch: torch.Tensor, patches_dim, rotation: bool):
    """Validation metric of ``spatial_diffusion.py:783-856`` / ``:921-950`` for a WHOLE batch.

    The reference loops over the puzzles on the host and runs two ``greedy_cost_assignment`` calls per puzzle (each an
    O(n) host-synchronising loop); here the predicted and the ground-truth assignments of every puzzle are ONE
    ``da_greedy_cost_assignment`` launch each.  Returns ``(correct [B] bool, piece_accuracy [N] bool)``:
    a piece counts when it lands on the same grid cell as in the ground truth (and, with ``rotation``, when the cosine
    between predicted and true rotation vectors exceeds cos(pi/4)); a puzzle is correct when all its pieces are."""
    dev = img.device
    B = int(batch.max()) + 1
    counts = torch.bincount(batch, minlength=B)
    ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(counts, 0)
    dims = patches_dim.tolist() if torch.is_tensor(patches_dim) else [list(d) for d in patches_dim]
    grid = torch.cat([real_grid(int(r), int(c), dev) for r, c in dims])
    if grid.shape[0] != img.shape[0]:
        raise ValueError("patches_dim does not match the number of pieces per puzzle")
    gp = ptr.to(torch.int32)
    ass_gt = greedy_cost_assignment_batched(x_gt[:, :2], grid, gp)
    ass_pr = greedy_cost_assignment_batched(img[:, :2], grid, gp)
    puzzle_of = batch.to(torch.int64)
    base = ptr[:-1][puzzle_of]                       # the kernel's indices are local to each puzzle, rows in greedy order

    def cell_of_piece(ass):                          # == ass[sort(ass[:, 0])][:, 1] per puzzle (:796-801)
        cell = torch.empty(img.shape[0], dtype=torch.int64, device=dev)
        cell[base + ass[:, 0]] = ass[:, 1]
        return cell

    piece_ok = cell_of_piece(ass_gt) == cell_of_piece(ass_pr)
    if rotation:
        import math

        rot_ok = torch.cosine_similarity(img[:, 2:], x_gt[:, 2:]) > math.cos(math.pi / 4)
        piece_ok = piece_ok & rot_ok
    wrong = torch.zeros(B, dtype=torch.int64, device=dev).index_add_(0, puzzle_of, (~piece_ok).to(torch.int64))
    return wrong == 0, piece_ok
