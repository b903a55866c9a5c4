"""PCA building blocks (Jacobi eigensolver, Gram / covariance PCA) against numpy / sklearn (the reference calls
sklearn.decomposition.PCA, transforms/cell_feature.py:60-62,176-177)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("g", [7, 64, 301])
def test_sym_eig_jacobi(cuda, g):
    from dance_b200 import ops
    rng = np.random.default_rng(g)
    A = rng.normal(size=(g + 20, g)).astype(np.float32)
    Cm = (A.T @ A).astype(np.float32)
    ev, V, sweeps = ops.sym_eig(torch.from_numpy(Cm.copy()).to(cuda))
    ev, V = ev.cpu().numpy().astype(np.float64), V.cpu().numpy().astype(np.float64)
    ref = np.linalg.eigvalsh(Cm.astype(np.float64))[::-1]
    assert np.allclose(ev, ref, rtol=2e-5, atol=1e-4 * ref[0])
    assert np.abs(V @ V.T - np.eye(g)).max() < 1e-4                       # orthonormal rows
    assert rel_err(V @ Cm.astype(np.float64) @ V.T, np.diag(ev)) < 1e-4   # diagonalises C
    assert sweeps <= 20


@pytest.mark.parametrize("n,f,k", [(500, 40, 10), (60, 300, 20), (1000, 200, 50)])
def test_pca_matches_sklearn(cuda, n, f, k):
    """explained variance, subspace and (sign-fixed) scores vs sklearn's exact solver; both the covariance side
    (n > f, CellPCA shape) and the Gram side (n < f, WeightedFeaturePCA gene-PCA shape)."""
    from sklearn.decomposition import PCA
    from dance_b200 import ops
    rng = np.random.default_rng(n + f)
    # decaying spectrum so that the leading components are well separated
    X = (rng.normal(size=(n, min(n, f))) * np.linspace(3, 0.2, min(n, f))) @ rng.normal(size=(min(n, f), f)) / np.sqrt(f) + rng.normal(size=f)
    X = X.astype(np.float32)
    ref = PCA(n_components=k, svd_solver="full")
    ref_scores = ref.fit_transform(X.astype(np.float64))
    # The reference pins scikit-learn==1.3.2 (requirements.txt:21) whose PCA uses the U-BASED svd_flip (largest-|.| entry
    # of every score column positive); the sklearn in this image (>=1.5) flips on Vt instead → convert to the pinned rule.
    flip = np.sign(ref_scores[np.abs(ref_scores).argmax(0), np.arange(k)])
    ref_scores = ref_scores * flip
    ref_components = ref.components_ * flip[:, None]
    out = ops.pca(torch.from_numpy(X).to(cuda), k)
    assert np.allclose(out["explained_variance"].cpu().numpy(), ref.explained_variance_, rtol=1e-3)
    comps = out["components"].cpu().numpy().astype(np.float64)
    cos = np.sum(comps * ref_components, axis=1)
    assert cos[:min(k, 8)].min() > 0.999                                  # leading directions agree, signs included
    scores = out["scores"].cpu().numpy().astype(np.float64)
    # the k-dimensional subspace: projecting the reference scores onto ours loses nothing
    q, _ = np.linalg.qr(scores)
    assert rel_err(q @ (q.T @ ref_scores), ref_scores) < 2e-3
    # sign convention (svd_flip, u-based): leading columns coincide with sklearn
    for c in range(min(k, 5)):
        assert rel_err(scores[:, c], ref_scores[:, c]) < 5e-3
