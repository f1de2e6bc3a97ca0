"""Configuration bag with the reference's attribute names and YAML layout (/root/reference/parser_sg.py:3-67,
/root/reference/config/config.yml).  `sgpr_args().load(path)` is what eval_pair.py / eval_batch.py / main_sg.py call.

Differences from the reference, on purpose: the YAML is read with an explicit SafeLoader (parser_sg.py:37 calls
`yaml.load` without a Loader, a TypeError on PyYAML >= 6), the file handle is closed, and a missing key keeps its
default instead of raising KeyError.
"""
import os

import yaml

# section -> attribute names read from it (same grouping as config/config.yml)
_SECTIONS = {
    "common": ("model", "cuda", "batch_size", "p_thresh", "graph_pairs_dir", "pair_list_dir"),
    "arch": ("keep_node", "filters_1", "filters_2", "filters_3", "tensor_neurons", "bottle_neck_neurons", "K",
             "knn_ties"),      # knn_ties: not a reference key — "cuda" (default) / "cpu", see INTEGRATION.md
    "train": ("epochs", "train_sequences", "eval_sequences", "dropout", "learning_rate", "weight_decay", "gpu",
              "logdir", "node_num"),
    "eva_batch": ("sequences", "output_path", "show"),
    "eva_pair": ("pair_file",),
}

_DEFAULTS = dict(
    model="", graph_pairs_dir="/dir_of_graph_pairs", p_thresh=3, batch_size=128, pair_list_dir="",
    keep_node=1, filters_1=64, filters_2=64, filters_3=32, tensor_neurons=16, bottle_neck_neurons=16, K=10,
    epochs=500, train_sequences=[], eval_sequences=[], dropout=0, learning_rate=1e-3, weight_decay=5e-4, gpu=0,
    logdir="./logs", node_num=100, sequences=[], output_path="./eva", show=False, pair_file="",
    knn_ties="cuda",
)


class sgpr_args():
    def __init__(self):
        for name, value in _DEFAULTS.items():
            setattr(self, name, list(value) if isinstance(value, list) else value)

    def load(self, config_file):
        with open(os.path.abspath(config_file)) as handle:
            tree = yaml.load(handle, Loader=yaml.SafeLoader) or {}
        for section, names in _SECTIONS.items():
            block = tree.get(section) or {}
            for name in names:
                if name in block:
                    setattr(self, name, block[name])
        if not hasattr(self, "cuda"):      # the reference only defines .cuda inside load() (parser_sg.py:58)
            self.cuda = str(self.gpu)
        return self
