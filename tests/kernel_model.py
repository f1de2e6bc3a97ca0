"""numpy model of the CUDA kernel's ARITHMETIC FORM (csrc/embed_kernel.cuh), driven by the packed weight blob the
C library produces on the host.  Test infrastructure: it lets the CPU suite check the packing rules (BN fold, sign
fold, Wa/Wb split, transposes) and the exactness of the EdgeConv refactor against the oracle without a GPU.
It is not a product path and is never imported by sg_pr_b200."""
import numpy as np


def _topk_set_lowest_index(pd, k):
    """Indices of the k largest per row; ties at the threshold resolved to the lowest indices (kernel rule)."""
    n = pd.shape[-1]
    order = np.lexsort((np.broadcast_to(np.arange(n), pd.shape), -pd), axis=-1)
    return np.sort(order[..., :k], axis=-1)


def _pd(x):
    """x [N, C] node-major -> pd [N, N] in the kernel's order: (2*dot - xx_j) - xx_i."""
    x = x.astype(np.float32)
    dot = (x @ x.T).astype(np.float32)
    xx = np.sum((x * x).astype(np.float32), axis=1, dtype=np.float32)
    return ((np.float32(2.0) * dot - xx[None, :]).astype(np.float32) - xx[:, None]).astype(np.float32)


def _lrelu(z):
    return np.where(z > 0, z, z * np.float32(0.2)).astype(np.float32)


def unpack_pairs(flat, cin, co):
    """Inverse of csrc/pack.hpp::pair_index: flat pair-layout buffer -> dense [cin, co] matrix."""
    cpl = co // 32
    ci = np.arange(cin)[:, None]
    c = np.arange(co)[None, :]
    lane, j = c // cpl, c % cpl
    idx = (ci >> 1) * 2 * co + (j >> 1) * 128 + lane * (4 if cpl >= 2 else 2) + (j & 1) * 2 + (ci & 1)
    return flat[idx]


def embed_graph(feat, k, blob, offs):
    """feat [15, N] -> dict(emb [N,32], att [N], pooled [32], knn [6][N,k], layers [6][N,C'])."""
    f32 = np.float32
    n = feat.shape[1]
    sec = lambda name, count: blob[offs[name]:offs[name] + count]
    out = {"knn": [], "layers": []}

    def edge_layer(x, wname, abname, cin, cout):
        idx = _topk_set_lowest_index(_pd(x), k)
        w = unpack_pairs(sec(wname, cin * 2 * cout), cin, 2 * cout)
        y = (x.astype(f32) @ w).astype(f32)
        a, b = y[:, :cout], y[:, cout:]
        m = a[idx].max(axis=1)
        ab = sec(abname, 2 * cout)
        z = _lrelu(((m - a) + b).astype(f32) * ab[:cout] + ab[cout:])
        out["knn"].append(idx)
        out["layers"].append(z)
        return z

    # xyz layer 1, direct form
    xyz = feat[:3].T.astype(f32)
    idx = _topk_set_lowest_index(_pd(xyz), k)
    s1 = sec("s1", 64 * 8).reshape(64, 8)
    d = xyz[idx] - xyz[:, None, :]                                # [N, k, 3]
    e = np.einsum("nkc,oc->nko", d, s1[:, :3]).astype(f32).max(axis=1)
    y = (e + xyz @ s1[:, 3:6].T).astype(f32)
    x = _lrelu(y * s1[:, 6] + s1[:, 7])
    out["knn"].append(idx)
    out["layers"].append(x)
    x = edge_layer(x, "w_s2", "ab_s2", 64, 64)
    xyz3 = edge_layer(x, "w_s3", "ab_s3", 64, 32)

    sem = feat[3:].T.astype(f32)
    x = edge_layer(sem, "w_f1", "ab_f1", 12, 64)
    x = edge_layer(x, "w_f2", "ab_f2", 64, 64)
    sem3 = edge_layer(x, "w_f3", "ab_f3", 64, 32)

    cat = np.concatenate([xyz3, sem3], axis=1)
    ab = sec("ab_end", 64)
    emb = _lrelu((cat @ unpack_pairs(sec("w_end", 64 * 32), 64, 32)).astype(f32) * ab[:32] + ab[32:])
    watt = sec("att_w", 1024).reshape(32, 32)
    ctx = np.tanh((emb @ watt).astype(f32).sum(axis=0, dtype=f32) / f32(n)).astype(f32)
    att = (1.0 / (1.0 + np.exp(-(emb @ ctx).astype(f32)))).astype(f32)
    out.update(emb=emb, att=att, pooled=(emb.T @ att).astype(f32))
    return out


def pair_score(e1, e2, blob, head, offs):
    f32 = np.float32
    w = blob[offs["ntn_w"]:offs["ntn_w"] + 32 * 512].reshape(32, 32, 16)
    v = blob[offs["ntn_v"]:offs["ntn_v"] + 1024].reshape(16, 64)
    b = blob[offs["ntn_b"]:offs["ntn_b"] + 16]
    s = np.einsum("a,abt,b->t", e1, w, e2).astype(f32)
    s = np.maximum(s + v @ np.concatenate([e1, e2]) + b, 0).astype(f32)
    h = np.maximum(head[:256].reshape(16, 16) @ s + head[256:272], 0).astype(f32)
    z = f32(head[272:288] @ h + head[288])
    return f32(1.0 / (1.0 + np.exp(-z)))
