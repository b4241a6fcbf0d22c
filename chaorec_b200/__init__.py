"""chaorec_b200: B200-native hot path of ChaoRec (propagation, full-rank evaluation, kNN graph).

The package is a thin host layer over ``libchaorec_b200.so`` (hand-written sm_100a CUDA behind the
C ABI of ``include/chaorec_b200.h``).  It has no CPU or eager-PyTorch fallback: every compute
entry point raises if the library has not been built or the inputs are not CUDA tensors.
"""
from . import _lib
from .propagate import (CSRGraph, LightGCNConv, bipartite_graph, csr_from_coo, csr_from_sparse_coo,
                        graph_from_edge_index, lightgcn_forward, propagate_mean, resample_masked_graph, spmm)
from .ranking import (MaskCSR, PositivesCSR, build_mask_csr, build_positives_csr, evaluate, evaluate_full,
                      gene_ranklist, hit_masks, mask_csr_from_edges, metrics_from_hits, rank_topk, rank_topk_sum)
from .train import GraphedLightGCNStep, bpr_loss, lightgcn_loss, sample_negatives, training_batches
from .cooccur import user_cooccurrence_topk, user_graph_dict
from .item_graph import ItemGraph, KnnPattern, knn_graph, normalized_values, original_graph, pattern_matmul
from .knn import (build_knn_neighbourhood, get_knn_adj_mat, knn_laplacian_ones, knn_self_excluded,
                  knn_sym_laplacian, knn_topk, l2_normalize)

__all__ = [
    "CSRGraph", "LightGCNConv", "bipartite_graph", "resample_masked_graph", "csr_from_coo", "csr_from_sparse_coo", "graph_from_edge_index",
    "lightgcn_forward", "propagate_mean", "spmm",
    "MaskCSR", "PositivesCSR", "build_mask_csr", "build_positives_csr", "evaluate", "evaluate_full",
    "gene_ranklist", "hit_masks", "mask_csr_from_edges", "metrics_from_hits", "rank_topk", "rank_topk_sum",
    "GraphedLightGCNStep", "bpr_loss", "lightgcn_loss", "sample_negatives", "training_batches",
    "build_knn_neighbourhood", "get_knn_adj_mat", "knn_laplacian_ones", "knn_self_excluded",
    "knn_sym_laplacian", "knn_topk", "l2_normalize",
    "user_cooccurrence_topk", "user_graph_dict",
    "ItemGraph", "KnnPattern", "knn_graph", "normalized_values", "original_graph", "pattern_matmul",
]
