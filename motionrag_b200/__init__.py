"""motionrag_b200 — B200-native motion-retrieval hot path of MCG-NJU/MotionRAG.

Text-embedding similarity search over the RAG table -> top-k -> gather of the retrieved
clips' motion features into the CAMA context tensor, as hand-written sm_100a CUDA kernels
behind a C ABI (include/mrag.h, libmrag.so), with the reference's own Python interface on
top (`RAGDatabase`, src/data/rag.py; the ActionTransformer context contract,
src/projects/condition/module.py:298-301). No CPU fallback exists anywhere in this package.
"""
from ._cabi import MragError, launch_count  # noqa: F401
from .cama import CamaTransformer  # noqa: F401
from .context import (MotionContext, attach, block_causal_mask, gather_context, select_refs,  # noqa: F401
                      sinusoid_table)
from .features import FeatureTableWriter, build_feature_table, load_feature_rows  # noqa: F401
from .parallel import (PeerExchange, ShardedRetriever, alloc_feature_block, open_peer_tables,  # noqa: F401
                       shard_range)
from .rag import RAGDatabase, save_table  # noqa: F401
from .store import EmbeddingStore, FeatureTable, SearchResult, merge_topk  # noqa: F401

__version__ = "0.1.0"
