"""The path as stock PyTorch ops on whatever device the module lives on — the form the reference itself executes
(/root/reference/sg_net.py:79-138, dgcnn.py:23-49, layers_batch.py) with autograd available.

BASELINE / TEST USE ONLY: tools/perf_probe.py, tools/train_bench.py and tools/parity_report.py time or compare against
it ("what the unmodified module code costs on the same GPU"), tests/test_host_logic.py checks it against the oracle.
Nothing in the product calls it: SG.forward runs the CUDA library in eval and in train mode.
"""
import torch

from . import dgcnn


def _edge_layer(model, x, block):
    return block(dgcnn.get_graph_feature(x, k=model.args.K)).max(dim=-1)[0]                     # sg_net.py:84-86


def dgcnn_conv_pass(model, x):
    xyz, sem = x[:, :3, :], x[:, 3:, :]
    for block in (model.dgcnn_s_conv1, model.dgcnn_s_conv2, model.dgcnn_s_conv3):
        xyz = _edge_layer(model, xyz, block)
    for block in (model.dgcnn_f_conv1, model.dgcnn_f_conv2, model.dgcnn_f_conv3):
        sem = _edge_layer(model, sem, block)
    return model.dgcnn_conv_end(torch.cat((xyz, sem), dim=1)).permute(0, 2, 1)


def forward_torch(model, f1, f2):
    """(score [B], att_1 [B,N,1], att_2 [B,N,1]) of an `SG` module by stock PyTorch ops (honours model.training)."""
    e1, e2 = dgcnn_conv_pass(model, f1), dgcnn_conv_pass(model, f2)
    p1, a1 = model.attention(e1)
    p2, a2 = model.attention(e2)
    s = model.tensor_network(p1, p2).permute(0, 2, 1)
    s = torch.nn.functional.relu(model.fully_connected_first(s))
    return torch.sigmoid(model.scoring_layer(s)).reshape(-1), a1, a2
