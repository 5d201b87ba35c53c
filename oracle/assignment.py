"""Oracle restatement of ``greedy_cost_assignment`` (test infrastructure).

Follows ``puzzle_diff/model/spatial_diffusion.py:179-216`` step for step (masked global minimum, row-major
first occurrence, retire the row and the column), in plain eager torch on the CPU."""
import torch


def greedy_cost_assignment_ref(pos1: torch.Tensor, pos2: torch.Tensor, separately_rounded: bool = False) -> torch.Tensor:
    """``separately_rounded=False`` is the reference verbatim (``torch.norm`` on the broadcast difference).
    torch's CPU norm kernel is not a plain sqrt(dx*dx + dy*dy) (it differs from it, from the fused-multiply-add
    form and from the CUDA norm kernel by 1 ulp on ~1 % of the entries), so for bit-exact comparisons of TIES the
    CUDA kernel's arithmetic -- every operation rounded separately in fp32 -- can be selected instead; on inputs
    without 1-ulp near-ties both give the same assignment."""
    if separately_rounded:
        d = pos1[:, None, :2] - pos2[None, :, :2]
        dist = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).sqrt()
    else:
        dist = torch.norm(pos1[:, None] - pos2, dim=2)
    assignments = torch.zeros(dist.size(0), 3, dtype=torch.int64)
    mask = torch.ones_like(dist, dtype=torch.bool)
    counter = 0
    while mask.sum() > 0:
        min_val, min_idx = dist[mask].min(dim=0)
        idx = int(min_idx.item())
        ret = mask.nonzero()[idx, :]
        i, j = ret[0], ret[1]
        assignments[counter, 0] = i
        assignments[counter, 1] = j
        assignments[counter, 2] = min_val
        counter += 1
        mask[i, :] = 0
        mask[:, j] = 0
    return assignments[:counter]


def puzzle_accuracy_ref(img, x_gt, batch, patches_dim, rotation):
    """The per-puzzle metric loop of ``spatial_diffusion.py:783-856`` (test infrastructure): returns
    ``(correct [B] bool, piece_accuracy [N] bool)``."""
    import math

    correct, piece = [], []
    for i in range(int(batch.max()) + 1):
        idx = torch.where(batch == i)[0]
        gt_pos, pos = x_gt[idx, :2], img[idx, :2]
        n_patches = patches_dim[i].tolist() if torch.is_tensor(patches_dim) else list(patches_dim[i])
        y = torch.linspace(-1, 1, n_patches[0])
        x = torch.linspace(-1, 1, n_patches[1])
        xy = torch.stack(torch.meshgrid(x, y, indexing="xy"), -1)
        real_grid = xy.reshape(-1, 2)                      # einops "x y c -> (x y) c"
        gt_ass = greedy_cost_assignment_ref(gt_pos, real_grid)
        gt_ass = gt_ass[torch.sort(gt_ass[:, 0])[1]]
        pred_ass = greedy_cost_assignment_ref(pos, real_grid)
        pred = pred_ass[torch.sort(pred_ass[:, 0])[1]][:, 1]
        ok = gt_ass[:, 1] == pred
        c = bool(ok.all())
        if rotation:
            rot_ok = torch.cosine_similarity(img[idx, 2:], x_gt[idx, 2:]) > math.cos(math.pi / 4)
            c = c and bool(rot_ok.all())
            ok = rot_ok * ok
        correct.append(c)
        piece.append(ok)
    return torch.tensor(correct), torch.cat(piece)
