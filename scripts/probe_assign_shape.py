"""Run the tcgen05 assignment a few times at one shape (for ncu):  python scripts/probe_assign_shape.py B nb M"""
import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from tests.test_gpu_assign_tc import _assign
B, nb, M = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, nb * 4, generator=gen, device=dev); g = torch.randn(B, nb * 4, generator=gen, device=dev)
E = torch.randn(nb, M, 8, generator=gen, device=dev)
for _ in range(3):
    _assign(x, g, E, M, 4, 4, 1, False)
torch.cuda.synchronize()
