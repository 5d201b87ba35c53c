"""Oracle restatement of the SO(3) helpers on the 3D path (test infrastructure).

* ``vec2skew`` / ``skew2vec`` / ``skew_to_rmat`` / ``log_rmat`` / ``so3_scale``
  follow ``puzzle_diff/model/utils_3d.py:991-1071`` and
  ``puzzle_diff/model/backbones/efficient_gat_3d.py:30-45``.
* ``matrix_to_quaternion`` / ``quaternion_to_matrix`` restate
  ``pytorch3d.transforms`` (not listed in any reference env file, absent here):
  real-first (w, x, y, z); 4-candidate ``sqrt_positive_part`` method with the 0.1
  floor.  Older pytorch3d releases do not standardise the sign, newer ones do, so
  parity on quaternions is always checked up to sign.
"""
import torch


def vec2skew(vec: torch.Tensor) -> torch.Tensor:
    skew = torch.repeat_interleave(torch.zeros_like(vec).unsqueeze(-1), 3, dim=-1)
    skew[..., 2, 1] = vec[..., 0]
    skew[..., 2, 0] = -vec[..., 1]
    skew[..., 1, 0] = vec[..., 2]
    return skew - skew.transpose(-1, -2)


def skew2vec(skew: torch.Tensor) -> torch.Tensor:
    vec = torch.zeros_like(skew[..., 0])
    vec[..., 0] = skew[..., 2, 1]
    vec[..., 1] = -skew[..., 2, 0]
    vec[..., 2] = skew[..., 1, 0]
    return vec


def skew_to_rmat(vmat: torch.Tensor) -> torch.Tensor:
    return torch.matrix_exp(vec2skew(vmat))


def log_rmat(r_mat: torch.Tensor) -> torch.Tensor:
    # utils_3d.py:1018-1046
    skew_mat = r_mat - r_mat.transpose(-1, -2)
    sk_vec = skew2vec(skew_mat)
    s_angle = sk_vec.norm(p=2, dim=-1) / 2
    c_angle = (torch.einsum("...ii", r_mat) - 1) / 2
    angle = torch.atan2(s_angle, c_angle)
    scale = angle / (2 * s_angle)
    scale[angle == 0.0] = 0.0
    log_r_mat = scale[..., None, None] * skew_mat
    nanlocs = log_r_mat[..., 0, 0].isnan()
    if bool(nanlocs.any()):
        nanmats = r_mat[nanlocs]
        _, eigvec = torch.linalg.eigh(nanmats)
        nan_axes = eigvec[..., -1, :]
        nan_angle = angle[nanlocs]
        log_r_mat[nanlocs] = vec2skew(nan_angle[..., None] * nan_axes)
    return log_r_mat


def so3_scale(rmat: torch.Tensor, scalars: torch.Tensor) -> torch.Tensor:
    # utils_3d.py:1049-1061
    logs = log_rmat(rmat)
    return torch.matrix_exp(logs * scalars[..., None, None])


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    ret = torch.zeros_like(x)
    positive = x > 0
    ret[positive] = torch.sqrt(x[positive])
    return ret


def matrix_to_quaternion(matrix: torch.Tensor, standardize: bool = False) -> torch.Tensor:
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(
        torch.stack(
            [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22],
            dim=-1,
        )
    )
    quat_by_rijk = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    best = torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    out = quat_candidates[best, :].reshape(batch_dim + (4,))
    if standardize:
        out = torch.where(out[..., 0:1] < 0, -out, out)
    return out


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            1 - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(quaternions.shape[:-1] + (3, 3))
