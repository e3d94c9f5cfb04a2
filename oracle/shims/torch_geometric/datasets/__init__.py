def _unavailable(*a, **k):
    raise RuntimeError("oracle shim: PyG datasets are not available offline")


Flickr = Yelp = PPI = Reddit = GNNBenchmarkDataset = _unavailable
