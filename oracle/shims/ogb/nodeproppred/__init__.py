class PygNodePropPredDataset:  # data only; unavailable offline
    def __init__(self, *a, **k):
        raise RuntimeError("ogb datasets are not available offline (oracle shim)")


class Evaluator:
    def __init__(self, *a, **k):
        raise RuntimeError("ogb evaluators are not available offline (oracle shim)")
