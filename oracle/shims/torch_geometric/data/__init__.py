class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Batch(Data):
    @staticmethod
    def from_data_list(lst):
        raise RuntimeError("oracle shim: datasets are not available offline")
