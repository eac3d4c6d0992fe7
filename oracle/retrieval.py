"""Oracle: video-retrieval similarity + top-k recall.  TEST INFRASTRUCTURE ONLY.

Restates tools/video_retrieval.py:174-197: optional L2 normalisation, sklearn `cosine_distances`
(= 1 - normalised dot product), full-row `np.argsort`, and for k in [1, 5, 10, 20, 50] a hit when
the query label appears among the labels of the k nearest gallery rows.
"""
import numpy as np

KS = (1, 5, 10, 20, 50)


def cosine_topk(queries: np.ndarray, gallery: np.ndarray, k: int):
    """Indices [Nq, k] of the k most cosine-similar gallery rows, most similar first, and their
    similarities.  Ties are broken towards the lower gallery index (stable argsort of distance)."""
    qn = queries / np.maximum(np.linalg.norm(queries, axis=1, keepdims=True), 1e-12)
    gn = gallery / np.maximum(np.linalg.norm(gallery, axis=1, keepdims=True), 1e-12)
    sim = qn.astype(np.float64) @ gn.astype(np.float64).T
    idx = np.argsort(-sim, axis=1, kind="stable")[:, :k]
    return idx, np.take_along_axis(sim, idx, axis=1)


def recall_hits(topk_idx: np.ndarray, query_labels: np.ndarray, gallery_labels: np.ndarray, ks=KS):
    """video_retrieval.py:192-197: {k: number of queries whose label is among the top-k labels}."""
    hits = {}
    for k in ks:
        lab = gallery_labels[topk_idx[:, :k]]
        hits[int(k)] = int((lab == query_labels[:, None]).any(axis=1).sum())
    return hits
