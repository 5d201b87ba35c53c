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
