"""Seeded synthetic inputs shared by CPU and GPU tests (SURVEY.md §8d geometry)."""
import numpy as np


def clustered_pairs(n, d, n_clusters=50, seed=0, noise_frac=0.0, dup_text_classes=0):
    """Clustered unit-norm image/text embeddings mimicking CLIP geometry.
    noise_frac: fraction of captions replaced by another caption of the same cluster
    (exact duplicate text rows -> ties).  dup_text_classes>0: text side drawn from only
    that many distinct vectors (classification datasets); returns label ids too."""
    rng = np.random.RandomState(seed)
    c = rng.standard_normal((n_clusters, d))
    c2 = rng.standard_normal((n_clusters, d))
    z = rng.randint(0, n_clusters, n)
    x = c[z] + 0.6 * rng.standard_normal((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if dup_text_classes:
        protos = rng.standard_normal((dup_text_classes, d))
        lab = z % dup_text_classes
        y = protos[lab]
    else:
        lab = None
        y = 0.5 * x * np.sqrt(d) / 1.0 + 0.5 * c2[z] + 0.6 * rng.standard_normal((n, d))
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    mislabel = np.zeros(n, bool)
    if noise_frac > 0:
        idx = rng.choice(n, int(noise_frac * n), replace=False)
        for i in idx:
            peers = np.nonzero(z == z[i])[0]
            peers = peers[peers != i]
            if len(peers):
                y[i] = y[rng.choice(peers)]
                mislabel[i] = True
    return x.astype(np.float32), y.astype(np.float32), lab, mislabel


def iid_pairs(n, d, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((n, d)); x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = rng.standard_normal((n, d)); y /= np.linalg.norm(y, axis=1, keepdims=True)
    return x.astype(np.float32), y.astype(np.float32)
