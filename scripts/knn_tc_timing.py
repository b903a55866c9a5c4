import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from dance_b200 import ops
import benchmarks.micro as m
dev = torch.device("cuda:0")
nn, d = 100000, int(sys.argv[1])
X = torch.randn(nn, d, device=dev) + torch.randn(10, d, device=dev)[torch.randint(0, 10, (nn, ), device=dev)] * 3
med, best = m.timeit(lambda: ops.knn(X, 15, return_dist=False), iters=3, warmup=1)
print("MS", med)
'''
for d in (128, 50):
    for name, env in (("tc", {}), ("simt", {"B2_KNN_NO_TC": "1"}), ("tc no-select", {"B2_KNN_TC_DEBUG": "2"}),
                      ("tc no-insert", {"B2_KNN_TC_DEBUG": "4"}), ("tc none", {"B2_KNN_TC_DEBUG": "3"})):
        out = subprocess.run([sys.executable, "-c", CHILD, str(d)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        print(d, name, [l for l in out.stdout.splitlines() if l.startswith("MS")] or out.stderr[-300:], flush=True)
