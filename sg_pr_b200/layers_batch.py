"""Parameter holders for the attention pooling and the neural tensor network, with the reference's class and
parameter names (`AttentionModule.weight_matrix`; `TenorNetworkModule` (sic) `.weight_matrix`,
`.weight_matrix_block`, `.bias` — /root/reference/layers_batch.py:3-83) so checkpoints load unchanged.

In eval AND in train mode the arithmetic of both modules runs in the CUDA library (csrc/embed_kernel.cuh,
csrc/train_kernels.cuh); the `forward` methods below are stock PyTorch ops kept only for `torch_baseline.py` — the
"unmodified module code on the same GPU" baseline of the benchmarks and the CUDA-reference of the tie-rule tests.
"""
import torch


class AttentionModule(torch.nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.weight_matrix = torch.nn.Parameter(torch.empty(args.filters_3, args.filters_3))
        torch.nn.init.xavier_uniform_(self.weight_matrix)

    def forward(self, embedding):
        # layers_batch.py:35-38: context = tanh(mean_n(E W)); a = sigmoid(E context); pooled = E^T a
        context = torch.tanh((embedding @ self.weight_matrix).mean(dim=1))
        scores = torch.sigmoid(embedding @ context.unsqueeze(-1))
        return embedding.transpose(1, 2) @ scores, scores


class TenorNetworkModule(torch.nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        f, t = args.filters_3, args.tensor_neurons
        self.weight_matrix = torch.nn.Parameter(torch.empty(f, f, t))
        self.weight_matrix_block = torch.nn.Parameter(torch.empty(t, 2 * f))
        self.bias = torch.nn.Parameter(torch.empty(t, 1))
        for p in (self.weight_matrix, self.weight_matrix_block, self.bias):
            torch.nn.init.xavier_uniform_(p)

    def forward(self, embedding_1, embedding_2):
        # layers_batch.py:78-82: s_t = e1^T W[:, :, t] e2 ; relu(s + V [e1; e2] + b)
        bilinear = torch.einsum("ba,act,bc->bt", embedding_1.squeeze(-1), self.weight_matrix,
                                embedding_2.squeeze(-1)).unsqueeze(-1)
        block = self.weight_matrix_block @ torch.cat((embedding_1, embedding_2), dim=1)
        return torch.relu(bilinear + block + self.bias)
