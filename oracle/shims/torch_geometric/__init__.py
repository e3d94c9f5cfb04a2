"""Pure-torch stand-in for the subset of `torch_geometric` (PyG 1.7.2 behaviour) that the
VQ-GNN reference imports.  TEST INFRASTRUCTURE ONLY (oracle/); see oracle/README.md."""
__version__ = "1.7.2+oracle-shim"
