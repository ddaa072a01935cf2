"""multimodalsimilar_b200 -- the ArcFace head hot path of forrestsocool/MultimodalSimilar, B200-native.

Public surface (mirrors /root/reference/arcface.py):
    ArcMarginProduct         drop-in nn.Module (single GPU)
    ShardedArcMarginProduct  the same head class-sharded over a process group (PartialFC-style)
    ArcFaceCEFunction        the autograd.Function behind both
    ops                      tensor-level wrappers over the C ABI (include/arcface_b200.h)
    CosineIndex, cosine_topk fused cosine top-k / faiss-style flat inner-product index (retrieval.py, K4)
    FusedHeadAdamW           AdamW for head weights that also emits the next forward's normalised rows (optim.py)
    install_reference_shim, load_reference_head, reference_state_dict   reference checkpoint compatibility
    two_stream_embed         fused cat(normalize(img), normalize(text)) in front of the two-stream model's head (prehead.py)
    MultiHeadArcFace         several heads on one embedding, one CUDA graph for all of them (prehead.py)

The compute lives in libarcface_b200.so (hand-written sm_100a CUDA: tcgen05 / TMEM / TMA); importing the
package does not load it, the first op does, and raises if it is missing -- there is no fallback path.
"""
from .head import ArcFaceCEFunction, ArcMarginProduct, FusedLogits  # noqa: F401
from .sharded import ShardedArcMarginProduct, shard_range  # noqa: F401
from .optim import FusedHeadAdamW  # noqa: F401
from .retrieval import CosineIndex, cosine_topk  # noqa: F401
from .checkpoint import install_reference_shim, load_reference_head, reference_state_dict  # noqa: F401
from .prehead import MultiHeadArcFace, TwoStreamConcat, two_stream_embed  # noqa: F401

__all__ = ["ArcMarginProduct", "ShardedArcMarginProduct", "ArcFaceCEFunction", "FusedLogits", "shard_range",
           "FusedHeadAdamW", "CosineIndex", "cosine_topk", "install_reference_shim", "load_reference_head", "reference_state_dict",
           "MultiHeadArcFace", "TwoStreamConcat", "two_stream_embed"]
__version__ = "0.1.0"
