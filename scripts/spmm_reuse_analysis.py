"""How much on-chip reuse could a tiled SpMM get on the benchmark's kNN graph?  (CPU analysis, numpy only.)
For tiles of 128 consecutive rows of Â, counts the distinct columns a tile touches vs its non-zeros, under three cell orders:
natural (random), sorted by cluster label, and sorted by nearest of 2048 random anchors within the cluster order."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
X = port.synthetic_embedding(n, d=128, n_clusters=10, seed=0)
adj, idx = port.feature2adj(X, 15)
A = (adj + __import__("scipy.sparse", fromlist=["eye"]).eye(n)).tocsr()


def reuse(order):
    inv = np.empty(n, dtype=np.int64)
    inv[order] = np.arange(n)
    P = A[order][:, order].tocsr()
    tot = uniq = 0
    for r0 in range(0, n, 128):
        cols = P.indices[P.indptr[r0]:P.indptr[min(n, r0 + 128)]]
        tot += cols.size
        uniq += np.unique(cols).size
    return tot / uniq


# cluster labels by k-means-free proxy: nearest of the 10 true centres is unknown here → use a cheap 10-means
rng = np.random.default_rng(0)
cent = X[rng.choice(n, 10, replace=False)]
for _ in range(10):
    lab = ((X[:, None, :] - cent[None]) ** 2).sum(-1).argmin(1)
    cent = np.stack([X[lab == c].mean(0) if (lab == c).any() else cent[c] for c in range(10)])
anch = X[rng.choice(n, 2048, replace=False)]
d2 = (X ** 2).sum(1)[:, None] + (anch ** 2).sum(1)[None] - 2 * X @ anch.T
near = d2.argmin(1)
print("n", n, "nnz/row", A.nnz / n)
print("gathers per distinct row in a 128-row tile:")
print("  natural order        %.3f" % reuse(np.arange(n)))
print("  cluster-sorted       %.3f" % reuse(np.argsort(lab, kind="stable")))
print("  anchor-sorted (2048) %.3f" % reuse(np.lexsort((near, lab))))
