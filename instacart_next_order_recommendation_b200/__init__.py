"""B200-native retrieval hot path of chen-bowen/instacart_next_order_recommendation.

Cosine scoring of query embeddings against the in-memory product catalog + top-k, the batched
IR-evaluator scoring, and MultipleNegativesRankingLoss — as hand-written sm_100a CUDA kernels
behind a C ABI (include/icr_b200.h), with Python drop-ins for the reference's call signatures.
Importing the package needs neither a GPU nor the built library; calling any compute entry
point does, and raises otherwise (there is no CPU fallback).
"""

from .evaluation import InformationRetrievalEvaluator, compute_ir_metrics, rank_all
from .index import DeviceCatalog, EmbeddingIndex
from .losses import MnrlStepGraph, MultipleNegativesRankingLoss, mnrl_loss, mnrl_loss_gathered, mnrl_step_graph
from .recommender import MonitoredRecommender, RecommendationMetrics, Recommender
from .sharded import ShardedCatalog, shard_bounds
from .similarity import cos_sim, cos_topk

__all__ = [
    "cos_sim", "cos_topk", "DeviceCatalog", "EmbeddingIndex", "Recommender", "MonitoredRecommender",
    "RecommendationMetrics", "MultipleNegativesRankingLoss", "MnrlStepGraph", "mnrl_step_graph", "mnrl_loss", "InformationRetrievalEvaluator",
    "rank_all", "compute_ir_metrics", "ShardedCatalog", "shard_bounds",
]
__version__ = "0.1.0"
