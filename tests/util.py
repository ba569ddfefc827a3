"""Seeded synthetic inputs shared by the CPU and GPU tests."""
import numpy as np


def gaussian(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)


def clustered_unit(n, d, seed, ncent=64, noise=0.3, centroids=None):
    """Dist U: Gaussian centroids + noise, rows L2-normalised."""
    rng = np.random.default_rng(seed)
    if centroids is None:
        centroids = np.random.default_rng(777).standard_normal((ncent, d), dtype=np.float32)
    x = centroids[rng.integers(0, centroids.shape[0], n)] + noise * rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def fingerprints(n, d, seed, p=0.05, dtype=np.int8):
    """Dist F: Morgan-bit-like 0/1 rows (retrieve/retrieve_faiss.py:36-44)."""
    return (np.random.default_rng(seed).random((n, d)) < p).astype(dtype)


def count_fingerprints(n, d, seed):
    """Difference-fingerprint-like small signed counts, int64 (retrieve/retrieve_faiss.py:18-27)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-2, 3, size=(n, d)) * (rng.random((n, d)) < 0.04)
    return x.astype(np.int64)
