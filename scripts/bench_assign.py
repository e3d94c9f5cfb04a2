import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from tests.test_gpu_assign_tc import _assign
dev = torch.device("cuda:0")
def timeit(B, nb, M, impl, with_stats, reps=5):
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, nb * 4, generator=gen, device=dev); g = torch.randn(B, nb * 4, generator=gen, device=dev)
    E = torch.randn(nb, M, 8, generator=gen, device=dev)
    _assign(x, g, E, M, 4, 4, impl, with_stats)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = []
    for _ in range(reps):
        torch.cuda.synchronize(); a.record(); _assign(x, g, E, M, 4, 4, impl, with_stats); b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    return min(t)
for (B, nb, M) in [(84663, 32, 256), (50000, 32, 1024), (6000, 151, 1024), (10000, 64, 4096), (20000, 32, 4096)]:
    print(B, nb, M, "tc+stats %.3f  tc-nostats %.3f  simt+stats %.3f  simt-nostats %.3f ms" % (
        timeit(B, nb, M, 1, True), timeit(B, nb, M, 1, False), timeit(B, nb, M, 0, True), timeit(B, nb, M, 0, False)))
