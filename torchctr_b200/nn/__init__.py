from .embedding import EmbeddingTable, PooledLookupGroup, SparseOptimizerBinding, pooled_lookup
from .dynamic import DynamicEmbedding
from .vocab import VocabIndex
from .interaction import CrossLayer, fm_interaction
from .linear import linear_tc, matmul_precision, set_matmul_precision

__all__ = ["EmbeddingTable", "DynamicEmbedding", "VocabIndex", "PooledLookupGroup", "SparseOptimizerBinding",
           "pooled_lookup", "CrossLayer", "fm_interaction", "linear_tc", "matmul_precision", "set_matmul_precision"]
