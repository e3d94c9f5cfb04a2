"""Load the UNMODIFIED VQ-GNN reference (`/root/reference/vq_gnn_v{1,2}`) on CPU.

TEST INFRASTRUCTURE ONLY (oracle/).  Only `tests/` and `oracle/gen_golden.py` may import this.
The reference needs torch_geometric / torch_sparse / torch_scatter / ogb, none of which are
installed; `oracle/shims/` provides pure-torch stand-ins for exactly the imported leaf symbols, so
the reference's own `vq.py`, `convs.py`, `models.py`, `utils/dataloader.py` (v1 `mapper`) and
`dataloader.py` (v2 `_k_hop_subgraph`) execute verbatim.  `/root/reference` exists only in the
builder container; on the GPU box `available()` is False and the committed fixtures under
`tests/golden/` (written by `oracle/gen_golden.py`) stand in.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root() -> str:
    """/root/reference in the builder container; on the GPU box the git-ignored copy `oracle/_ref/` that
    `make oracle_ref` made there (it travels with the gpurun snapshot like the built .so)."""
    for p in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(p, "vq_gnn_v2", "vq.py")):
            return p
    return "/root/reference"


REF_ROOT = os.environ.get("VQGNN_REFERENCE_ROOT") or _default_root()
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_CACHE = {}
_SHADOWED = ("vq", "convs", "models", "utils", "dataloader", "models_inductive")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "vq_gnn_v2", "vq.py"))


def load_reference(version: str = "v2") -> types.SimpleNamespace:
    """Return a namespace with the reference's modules for `version` in {"v1", "v2"}."""
    if version in _CACHE:
        return _CACHE[version]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    root = os.path.join(REF_ROOT, f"vq_gnn_{version}")
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                  if k in _SHADOWED or k.split(".")[0] == "utils"}
    saved_path = list(sys.path)
    sys.path.insert(0, root)
    if _SHIMS not in sys.path:
        sys.path.insert(1, _SHIMS)
    try:
        ns = types.SimpleNamespace(version=version, root=root)
        ns.vq = importlib.import_module("vq")
        ns.convs = importlib.import_module("convs")
        ns.models = importlib.import_module("models")
        if version == "v1":
            ns.mapper = importlib.import_module("utils.dataloader").mapper
            ns.dataloader = importlib.import_module("utils.dataloader")
        else:
            ns.dataloader = importlib.import_module("dataloader")
    finally:
        # un-shadow: keep the reference modules reachable only through `ns`
        for k in list(sys.modules):
            if k in _SHADOWED or k.split(".")[0] == "utils":
                sys.modules.pop(k)
        sys.modules.update(saved_mods)
        sys.path[:] = [p for p in saved_path]
        if _SHIMS not in sys.path:
            sys.path.append(_SHIMS)  # the loaded reference modules keep referring to the shims
    _CACHE[version] = ns
    return ns


def shim_sparse():
    """The shimmed `torch_sparse` module (so fixtures can build SparseTensor inputs)."""
    if _SHIMS not in sys.path:
        sys.path.append(_SHIMS)
    return importlib.import_module("torch_sparse")
