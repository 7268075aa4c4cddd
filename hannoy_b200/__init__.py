"""hannoy_b200 — B200-native batched HNSW search behind hannoy's Reader / QueryBuilder / Distance API.

Host-side mirror (Python) of the reference's public search surface (src/reader.rs, src/distance/),
calling the CUDA engine only through the C-ABI of libhannoy_b200.so (include/hannoy_b200.h).
"""
from .reader import (  # noqa: F401
    BinaryQuantizedCosine, BinaryQuantizedEuclidean, BinaryQuantizedManhattan, CancelToken, Cosine, Distance, Euclidean,
    Hamming, HannoyError, InvalidVecDimension, Manhattan, MissingMetadata, NeedBuild, QueryBuilder, Reader,
    Searched, UnmatchingDistance, exact_knn, merge_topk_device,
)

__all__ = [
    "Reader", "QueryBuilder", "Searched", "CancelToken", "Distance", "Euclidean", "Cosine", "Manhattan", "Hamming",
    "BinaryQuantizedCosine", "BinaryQuantizedEuclidean", "BinaryQuantizedManhattan", "HannoyError",
    "InvalidVecDimension", "MissingMetadata", "UnmatchingDistance", "NeedBuild", "exact_knn", "merge_topk_device",
]
