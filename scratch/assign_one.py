import sys; sys.path.insert(0, '/root/repo')
import torch
from tests.test_gpu_assign_tc import _assign
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
B, nb, M = 84663, 32, 256
x = torch.randn(B, nb * 4, generator=gen, device=dev); g = torch.randn(B, nb * 4, generator=gen, device=dev)
E = torch.randn(nb, M, 8, generator=gen, device=dev)
for _ in range(3):
    _assign(x, g, E, M, 4, 4, 1, True)
torch.cuda.synchronize()
