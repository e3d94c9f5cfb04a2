import sys, time; sys.path.insert(0, '/root/repo')
import torch, bench
from vq_gnn_b200.loader import DevicePrefetcher
dev = torch.device("cuda:0")
g, batches = bench.build_workload(dev, 0, 1, 1.0, n_batches=4)
model = bench.build_model(dev, g.N, False, 'auto')
opt = torch.optim.RMSprop(model.parameters(), lr=1e-3, alpha=0.99)
bench.warm_start(model, batches)
host = []
for x, bA, y in batches:
    pin = lambda t: None if t is None else (tuple(u.cpu().pin_memory() for u in t) if isinstance(t, tuple) else t.cpu().pin_memory())
    host.append((x.cpu().pin_memory(), tuple(pin(t) for t in bA), y.cpu().pin_memory()))
import threading
LOG = []
def prep(b):
    t1 = time.perf_counter()
    x, bA, y = b
    p = model.prepare(bA)
    t2 = time.perf_counter()
    LOG.append(f"  prepare(cpu) {1e3*(t2-t1):.2f}")
    return x, p, y
side = torch.cuda.Stream(device=dev)
def run(n, verbose):
    pf = DevicePrefetcher(host, dev, prepare=prep, count=n, stream=side)
    for i in range(n):
        t0 = time.perf_counter()
        x, plan, y = pf.next()
        t1 = time.perf_counter()
        loss = bench.train_step(model, opt, x, plan, y, False)
        t2 = time.perf_counter()
        l = float(loss.item())
        t3 = time.perf_counter()
        if verbose: print(f"  cudaMallocs so far {torch.cuda.memory_stats()['num_device_alloc']}", end=" ")
        if verbose: print(f"step {i}: wait {1e3*(t1-t0):.2f} launch {1e3*(t2-t1):.2f} sync {1e3*(t3-t2):.2f} ms")
    pf.drain()
run(5, False)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(10, True); torch.cuda.synchronize(); print("\n".join(LOG[-10:])); print("total per step", (time.perf_counter()-t0)*100, "ms")
