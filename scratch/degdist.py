import sys; sys.path.insert(0, '/root/repo')
import torch, bench
dev = torch.device("cuda:0")
g, batches = bench.build_workload(dev, 0, 1, 1.0, n_batches=4)
for x, bA, y in batches[1:3]:
    r = bA[1][0]
    deg = torch.bincount(r, minlength=6000).float()
    tot = deg.sum()
    print("rows", deg.numel(), "entries", int(tot), "max deg", int(deg.max()), "median", int(deg.median()))
    for th in (128, 256, 512, 1024, 2048, 4096, 8192, 16384):
        m = deg > th
        print(f"  deg > {th:5d}: rows {int(m.sum()):5d}  entries {float(deg[m].sum()/tot)*100:5.1f}%")
