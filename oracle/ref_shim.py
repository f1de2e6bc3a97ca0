"""Import the UNMODIFIED reference (/root/reference) on a CPU-only box.  TEST INFRASTRUCTURE ONLY.

Build-container use only (the GPU box has no /root/reference): `oracle/make_golden.py` uses this to run the
reference's own `SG.forward` and record golden vectors (`oracle/make_golden_train.py` likewise for two optimiser
steps); the reference's own scripts are executed separately by `tests/test_reference_scripts.py`, on top of the drop-in
modules.  No reference source is copied — the
reference modules are imported from where they lie, behind import-time stubs for what this image lacks
(texttable / tensorboardX / matplotlib), a no-op `.cuda()`, a CPU `torch.device` for dgcnn.py:32, and a
`Loader` for the PyYAML-6 `yaml.load` call at parser_sg.py:37.  Recipe: SURVEY.md §8(c).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SGPR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sg_net.py"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    for key, value in attrs.items():
        setattr(mod, key, value)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


class _Writer:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass


class _Table:
    def add_rows(self, rows):
        self.rows = rows

    def draw(self):
        return "\n".join(str(r) for r in getattr(self, "rows", []))


_LOADED = None


def load_reference():
    """Returns the imported reference modules as a namespace (sg_net, dgcnn, layers_batch, parser_sg, utils)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch
    import yaml

    _stub("texttable", Texttable=_Table)
    _stub("tensorboardX", SummaryWriter=_Writer)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    tk = _stub("mpl_toolkits")
    tk.mplot3d = _stub("mpl_toolkits.mplot3d", Axes3D=object)

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    real_yaml_load = yaml.load
    yaml.load = lambda stream, Loader=None: real_yaml_load(stream, Loader=Loader or yaml.SafeLoader)
    real_torch_load = torch.load
    dev = "cpu" if not torch.cuda.is_available() else None

    def _load(path, map_location=None, **kw):
        kw.setdefault("weights_only", False)
        return real_torch_load(path, map_location=dev or map_location, **kw)

    torch.load = _load

    # the reference's module names (sg_net, utils, ...) must win over anything same-named on sys.path
    for name in ("sg_net", "dgcnn", "layers_batch", "parser_sg", "utils"):
        sys.modules.pop(name, None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import dgcnn
        import layers_batch
        import parser_sg
        import utils
        import sg_net
    finally:
        sys.path.remove(REFERENCE_ROOT)

    if not torch.cuda.is_available():
        class _TorchProxy:
            def __getattr__(self, item):
                return getattr(torch, item)

            @staticmethod
            def device(*a, **k):
                return torch.device("cpu")

        dgcnn.torch = _TorchProxy()

    _LOADED = types.SimpleNamespace(sg_net=sg_net, dgcnn=dgcnn, layers_batch=layers_batch,
                                    parser_sg=parser_sg, utils=utils)
    # leave the names importable for the reference's own scripts, but do not shadow the product's modules
    for name in ("sg_net", "dgcnn", "layers_batch", "parser_sg", "utils"):
        sys.modules.pop(name, None)
    return _LOADED


def reference_trainer(K: int | None = None, node_num: int | None = None, model_path: str | None = None):
    """`SGTrainer(args, False)` of the reference with config/config.yml, optionally overriding K / node_num."""
    ref = load_reference()
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        args = ref.parser_sg.sgpr_args()
        args.load("./config/config.yml")
        if K is not None:
            args.K = K
        if node_num is not None:
            args.node_num = node_num
        if model_path is not None:
            args.model = model_path
        args.logdir = "/tmp/sgpr_ref_logs"
        trainer = ref.sg_net.SGTrainer(args, False)
    finally:
        os.chdir(cwd)
    trainer.model.eval()
    return trainer


def reference_module(trainer):
    return trainer.model.module if hasattr(trainer.model, "module") else trainer.model
