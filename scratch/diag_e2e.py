import sys, time; sys.path.insert(0, '/root/repo')
import torch, bench
dev = torch.device("cuda:0")
g, batches = bench.build_workload(dev, 0, 1, 1.0, n_batches=2)
model = bench.build_model(dev, g.N, False, 1)
x, bA, y = batches[1]
pin = lambda t: None if t is None else (tuple(u.cpu().pin_memory() for u in t) if isinstance(t, tuple) else t.cpu().pin_memory())
hx, hA, hy = x.cpu().pin_memory(), tuple(pin(t) for t in bA), y.cpu().pin_memory()
def to_dev(t):
    if t is None: return None
    if isinstance(t, tuple): return tuple(u.to(dev, non_blocking=True) for u in t)
    return t.to(dev, non_blocking=True)
def nbytes(t):
    if t is None: return 0
    if isinstance(t, tuple): return sum(nbytes(u) for u in t)
    return t.numel() * t.element_size()
nb = nbytes(hx) + nbytes(hA) + nbytes(hy)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dx = hx.to(dev, non_blocking=True); dA = tuple(to_dev(t) for t in hA)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    plan = model.prepare(dA)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    sp = plan.split_v1(); plan.chunk_rows('bwd')
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"H2D {nb/1e6:.0f} MB: {(t1-t0)*1e3:.2f} ms ({nb/(t1-t0)/1e9:.1f} GB/s); prepare {(t2-t1)*1e3:.2f} ms; split+chunks {(t3-t2)*1e3:.2f} ms")
for name, t in zip(["deg_inv","A_BN","A_BB","A_NB_v","batch_idx"], hA):
    print(name, nbytes(t)/1e6, "MB", (t[0].dtype if isinstance(t, tuple) else (t.dtype if t is not None else None)))
