class ToSparseTensor:  # data-prep only; never on the hot path
    def __init__(self, *a, **k):
        pass

    def __call__(self, data):
        raise RuntimeError("oracle shim: dataset transforms are not available offline")


class ToUndirected(ToSparseTensor):
    pass


class Compose(ToSparseTensor):
    pass
