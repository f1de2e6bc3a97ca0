"""Hand-derived forward/backward of one training step in the arithmetic form the CUDA training kernels use
(per-node A = X Wa^T, B = X Wb^T; y_ij = A_j - A_i + B_i; BatchNorm batch statistics and its backward folded into
per-node sums and two scatters).  Plain torch tensor algebra WITHOUT autograd: tests/test_train_form.py checks every
gradient against the oracle's autograd, so a formula error shows up on the CPU before any kernel exists.

Layout: a "side" is one BatchNorm batch (sg_net.py:123-124); x is node-major [G, N, C]; idx [G, N, k] int64.
"""
from __future__ import annotations

import torch

EPS = 1e-5
SLOPE = 0.2


def lrelu(z):
    return torch.where(z > 0, z, SLOPE * z)


def lrelu_grad(z):
    return torch.where(z > 0, torch.ones_like(z), torch.full_like(z, SLOPE))


def knn(x, k):
    """dgcnn.py:14-20 on node-major x [G, N, C]."""
    inner = -2 * torch.matmul(x, x.transpose(1, 2))
    xx = (x ** 2).sum(-1, keepdim=True)
    pd = -xx.transpose(1, 2) - inner - xx
    return pd.topk(k, dim=-1)[1]


def gather_rows(a, idx):
    """a [G, N, C], idx [G, N, k] -> [G, N, k, C]"""
    g, n, k = idx.shape
    return torch.gather(a.unsqueeze(1).expand(g, n, n, a.shape[-1]), 2, idx.unsqueeze(-1).expand(g, n, k, a.shape[-1]))


def edgeconv_fwd(x, idx, w, gamma, beta, direct=False):
    """Returns (out [G,N,C'], cache).  direct=True (xyz layer 1): the forward values come from the reference's own
    form W_a (x_j - x_i) + W_b x_i per edge — metre-scale coordinates lose ~5 bits in A_j - A_i; the backward keeps
    the A/B form, where that error only touches a correction term."""
    c = x.shape[-1]
    wa, wb = w[:, :c], w[:, c:]
    a, b = x @ wa.t(), x @ wb.t()
    d = b - a
    an = gather_rows(a, idx)                                  # [G,N,k,C']
    y = an + d.unsqueeze(2)
    if direct:
        xn = gather_rows(x, idx)
        y = (xn - x.unsqueeze(2)) @ wa.t() + (x @ wb.t()).unsqueeze(2)
    e = y.shape[0] * y.shape[1] * y.shape[2]
    mu = y.double().sum((0, 1, 2)) / e
    var = (y.double() ** 2).sum((0, 1, 2)) / e - mu ** 2
    mu, var = mu.float(), var.float()
    istd = torch.rsqrt(var + EPS)
    ymax, jmax = y.max(dim=2)
    ymin, jmin = y.min(dim=2)
    pos = gamma >= 0
    yext = torch.where(pos, ymax, ymin)
    jext = torch.where(pos, jmax, jmin)                        # position in the neighbour list
    yhat = (yext - mu) * istd
    z = gamma * yhat + beta
    cache = dict(x=x, idx=idx, w=w, a=a, d=d, sa=an.sum(2), mu=mu, var=var, istd=istd, yext=yext, jext=jext, z=z,
                 yhat=yhat, gamma=gamma, e=e)
    return lrelu(z), cache


def edgeconv_bwd(gout, c):
    """gout [G,N,C'] -> (dx [G,N,C], dw [C',2C], dgamma, dbeta)."""
    x, idx, w, a, d = c["x"], c["idx"], c["w"], c["a"], c["d"]
    g_, n, k = idx.shape
    cin = x.shape[-1]
    gz = gout * lrelu_grad(c["z"])
    dbeta = gz.sum((0, 1))
    dgamma = (gz * c["yhat"]).sum((0, 1))
    s = c["gamma"] * c["istd"]
    p = s * dbeta / c["e"]
    q = s * dgamma / c["e"] * c["istd"]
    r = p - q * c["mu"]
    # T_i = sum_j dy_ij
    sum_y = c["sa"] + k * d                                   # sum_j y_ij = SA - kA + kB
    t = s * gz - k * r - q * sum_y
    db = t
    # scatters: S1[n] = sum_{i: ext neighbour of (i,c) is n} gz[i,c];  S2[n] = sum_{i: n in N(i)} D[i];  deg[n]
    cout = gz.shape[-1]
    # neighbour NODE achieving the extreme: idx[g, i, jext[g,i,c]]
    node = torch.gather(idx.unsqueeze(-1).expand(g_, n, k, cout), 2, c["jext"].unsqueeze(2)).squeeze(2)   # [G,N,C']
    s1 = torch.zeros_like(gz).scatter_add_(1, node, gz)
    flat = idx.reshape(g_, n * k)
    s2 = torch.zeros_like(gz).scatter_add_(1, flat.unsqueeze(-1).expand(g_, n * k, cout),
                                           d.unsqueeze(2).expand(g_, n, k, cout).reshape(g_, n * k, cout))
    deg = torch.zeros(g_, n, dtype=gz.dtype).scatter_add_(1, flat, torch.ones(g_, n * k, dtype=gz.dtype))
    da = -t + s * s1 - deg.unsqueeze(-1) * (r + q * a) - q * s2
    wa, wb = w[:, :cin], w[:, cin:]
    dx = da @ wa + db @ wb
    dw = torch.cat([torch.einsum("gnc,gni->ci", da, x), torch.einsum("gnc,gni->ci", db, x)], dim=1)
    return dx, dw, dgamma, dbeta


def conv_end_fwd(xcat, w, gamma, beta):
    y = xcat @ w.t()
    e = y.shape[0] * y.shape[1]
    mu = (y.double().sum((0, 1)) / e)
    var = ((y.double() ** 2).sum((0, 1)) / e - mu ** 2).float()
    mu = mu.float()
    istd = torch.rsqrt(var + EPS)
    yhat = (y - mu) * istd
    z = gamma * yhat + beta
    return lrelu(z), dict(x=xcat, w=w, y=y, mu=mu, var=var, istd=istd, yhat=yhat, z=z, gamma=gamma, e=e)


def conv_end_bwd(gout, c):
    gz = gout * lrelu_grad(c["z"])
    dbeta, dgamma = gz.sum((0, 1)), (gz * c["yhat"]).sum((0, 1))
    s = c["gamma"] * c["istd"]
    dy = s * (gz - dbeta / c["e"] - c["yhat"] * dgamma / c["e"])
    return dy @ c["w"], torch.einsum("gnc,gni->ci", dy, c["x"]), dgamma, dbeta


def attention_fwd(emb, w):
    n = emb.shape[1]
    cbar = (emb @ w).mean(1)
    ctx = torch.tanh(cbar)
    att = torch.sigmoid(torch.einsum("gnf,gf->gn", emb, ctx))
    pooled = torch.einsum("gn,gnf->gf", att, emb)
    return pooled, dict(emb=emb, w=w, ctx=ctx, att=att, n=n)


def attention_bwd(dp, c):
    emb, att, ctx, w, n = c["emb"], c["att"], c["ctx"], c["w"], c["n"]
    de = att.unsqueeze(-1) * dp.unsqueeze(1)
    dsig = torch.einsum("gnf,gf->gn", emb, dp) * att * (1 - att)
    de = de + dsig.unsqueeze(-1) * ctx.unsqueeze(1)
    dctx = torch.einsum("gn,gnf->gf", dsig, emb)
    dcbar = dctx * (1 - ctx ** 2) / n
    de = de + (dcbar @ w.t()).unsqueeze(1)
    dw = torch.einsum("ga,gb->ab", emb.sum(1), dcbar)
    return de, dw


def head_fwd(e1, e2, ntn_w, ntn_v, ntn_b, w1, b1, w2, b2):
    s = torch.einsum("pa,abt,pb->pt", e1, ntn_w, e2) + torch.cat([e1, e2], 1) @ ntn_v.t() + ntn_b.reshape(1, -1)
    nt = torch.relu(s)
    hpre = nt @ w1.t() + b1
    h = torch.relu(hpre)
    zf = h @ w2.reshape(-1) + b2.reshape(())
    return torch.sigmoid(zf), dict(e1=e1, e2=e2, ntn_w=ntn_w, ntn_v=ntn_v, s=s, nt=nt, hpre=hpre, h=h, w1=w1, w2=w2)


def head_bwd(pred, target, c):
    bt = pred.shape[0]
    dzf = (pred - target) / bt
    dw2 = (dzf.unsqueeze(1) * c["h"]).sum(0).reshape(1, -1)
    db2 = dzf.sum().reshape(1)
    dh = dzf.unsqueeze(1) * c["w2"].reshape(1, -1) * (c["hpre"] > 0)
    dw1 = dh.t() @ c["nt"]
    db1 = dh.sum(0)
    ds = (dh @ c["w1"]) * (c["s"] > 0)
    e1, e2, w, v = c["e1"], c["e2"], c["ntn_w"], c["ntn_v"]
    dntn_w = torch.einsum("pa,pb,pt->abt", e1, e2, ds)
    dntn_v = ds.t() @ torch.cat([e1, e2], 1)
    dntn_b = ds.sum(0).reshape(-1, 1)
    f = e1.shape[1]
    de1 = torch.einsum("abt,pb,pt->pa", w, e2, ds) + ds @ v[:, :f]
    de2 = torch.einsum("abt,pa,pt->pb", w, e1, ds) + ds @ v[:, f:]
    return de1, de2, dict(ntn_w=dntn_w, ntn_v=dntn_v, ntn_b=dntn_b, w1=dw1, b1=db1, w2=dw2, b2=db2)


XYZ = ("dgcnn_s_conv1", "dgcnn_s_conv2", "dgcnn_s_conv3")
SEM = ("dgcnn_f_conv1", "dgcnn_f_conv2", "dgcnn_f_conv3")


def side_fwd(feat, k, sd, knn_override=None):
    """feat [G, 15, N] -> pooled [G, 32] + caches.  knn_override: the six index tensors to use instead of this model's
    own k-NN (isolates the arithmetic from near-tie neighbour flips, which in train mode perturb every output through
    the batch statistics)."""
    caches = {}
    li = 0
    xyz = feat[:, :3, :].transpose(1, 2).contiguous()
    sem = feat[:, 3:, :].transpose(1, 2).contiguous()
    for names, x, tag in ((XYZ, xyz, "xyz"), (SEM, sem, "sem")):
        for layer in names:
            idx = knn(x, k) if knn_override is None else knn_override[li]
            li += 1
            x, caches[layer] = edgeconv_fwd(x, idx, sd[layer + ".0.weight"].reshape(sd[layer + ".0.weight"].shape[0], -1),
                                            sd[layer + ".1.weight"], sd[layer + ".1.bias"],
                                            direct=(layer == "dgcnn_s_conv1"))
        caches[tag + "_out"] = x
    xcat = torch.cat([caches["xyz_out"], caches["sem_out"]], dim=-1)
    emb, caches["end"] = conv_end_fwd(xcat, sd["dgcnn_conv_end.0.weight"].reshape(32, 64), sd["dgcnn_conv_end.1.weight"],
                                      sd["dgcnn_conv_end.1.bias"])
    pooled, caches["att"] = attention_fwd(emb, sd["attention.weight_matrix"])
    return pooled, caches


def side_bwd(dp, caches, grads):
    de, dw = attention_bwd(dp, caches["att"])
    grads["attention.weight_matrix"] = grads.get("attention.weight_matrix", 0) + dw
    dxcat, dwend, dg, db = conv_end_bwd(de, caches["end"])
    acc(grads, "dgcnn_conv_end.0.weight", dwend.reshape(32, 64, 1))
    acc(grads, "dgcnn_conv_end.1.weight", dg)
    acc(grads, "dgcnn_conv_end.1.bias", db)
    for names, gout in ((XYZ, dxcat[..., :32]), (SEM, dxcat[..., 32:])):
        for layer in reversed(names):
            gout, dw_, dg, db = edgeconv_bwd(gout, caches[layer])
            acc(grads, layer + ".0.weight", dw_.reshape(dw_.shape[0], dw_.shape[1], 1, 1))
            acc(grads, layer + ".1.weight", dg)
            acc(grads, layer + ".1.bias", db)


def acc(grads, name, value):
    grads[name] = grads[name] + value if name in grads else value


def loss_and_grads(sd, f1, f2, target, k, knn_1=None, knn_2=None):
    """Mean-BCE loss, predictions and d loss / d every parameter, by the hand-derived backward."""
    p1, c1 = side_fwd(f1, k, sd, knn_1)
    p2, c2 = side_fwd(f2, k, sd, knn_2)
    pred, ch = head_fwd(p1, p2, sd["tensor_network.weight_matrix"], sd["tensor_network.weight_matrix_block"],
                        sd["tensor_network.bias"], sd["fully_connected_first.weight"], sd["fully_connected_first.bias"],
                        sd["scoring_layer.weight"], sd["scoring_layer.bias"])
    loss = torch.nn.functional.binary_cross_entropy(pred, target).mean()
    de1, de2, hg = head_bwd(pred, target, ch)
    grads = {"tensor_network.weight_matrix": hg["ntn_w"], "tensor_network.weight_matrix_block": hg["ntn_v"],
             "tensor_network.bias": hg["ntn_b"], "fully_connected_first.weight": hg["w1"],
             "fully_connected_first.bias": hg["b1"], "scoring_layer.weight": hg["w2"], "scoring_layer.bias": hg["b2"]}
    side_bwd(de1, c1, grads)
    side_bwd(de2, c2, grads)
    return float(loss), pred, grads
