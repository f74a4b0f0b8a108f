"""pixelrec_multimodal_b200 — B200 (sm_100a) full-catalogue scoring, top-K ranking
and ranking metrics behind the PixelRec_Multimodal inference API.  The arithmetic
lives in libpxr.so (csrc/, C ABI in include/pxr.h); this package is the Python
host that mirrors the reference interfaces."""
from .model import FastMultimodalRecommender  # noqa: F401
from .recommender import FastRecommender, ItemFeatureStore  # noqa: F401
from .evaluation import FullCatalogueEvaluator, RankingEvaluator, SampledRetrievalEvaluator, ranking_metrics  # noqa: F401
from .sharding import ShardedTopK, shard_range  # noqa: F401
from ._lib import PxrError  # noqa: F401

__all__ = ["FastMultimodalRecommender", "FastRecommender", "ItemFeatureStore", "FullCatalogueEvaluator", "SampledRetrievalEvaluator",
           "RankingEvaluator", "ranking_metrics", "ShardedTopK", "shard_range", "PxrError"]
