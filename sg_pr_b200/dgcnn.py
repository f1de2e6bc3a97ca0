"""`knn` / `get_graph_feature` with the reference's names and tensor contracts (/root/reference/dgcnn.py:14-49) as
differentiable stock PyTorch ops on whatever device `x` lives on.  Used by `torch_baseline.py` only (benchmark baseline,
tie-rule reference); the product fuses both into csrc/embed_kernel.cuh (eval) and csrc/train_kernels.cuh (training).  The DGCNN / PointNet classifier zoo of the reference file (dgcnn.py:52-149)
is dead code for SG_PR and is not provided."""
import torch


def knn(x, k):
    """x [B, C, N] -> idx [B, N, k]: k largest of  -|x_i - x_j|^2  per row (self included), dgcnn.py:14-20."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]


def get_graph_feature(x, k=20, cuda=0, idx=None, xyz=False):
    """x [B, C, N] -> edge tensor [B, 2C, N, k] = cat(x_j - x_i, x_i) over the k nearest j (dgcnn.py:23-49).
    `cuda` is accepted for signature compatibility; the device is taken from `x`."""
    b, c, n = x.shape
    if idx is None:
        idx = knn(x[:, :3, :] if xyz else x, k=k)
    rows = x.transpose(2, 1)
    nbr = rows[torch.arange(b, device=x.device)[:, None, None], idx]
    ctr = rows[:, :, None, :].expand(-1, -1, idx.shape[-1], -1)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)
