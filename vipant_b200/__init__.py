"""vipant_b200 -- B200-native InfoNCE / retrieval-scoring hot path of VIP-ANT (zhaoyanpeng/vipant).

Host side: Python mirror of the reference's loss-head API (``loss_head``), tensor-level functions
(``functional``) and the ctypes binding (``_cabi``) of the C-ABI CUDA library built from ``csrc/``;
``embed_cache``: packed shards of pre-computed embeddings (the on-disk format either side of the path);
``encoder_tail``: the towers' last LayerNorm + projection + normalisation fused up to the loss operands.
"""
from . import embed_cache  # noqa: F401
from .encoder_tail import FusedPostEncoder, encoder_tail  # noqa: F401
from .functional import infonce_loss, infonce_multi_loss, l2_normalize, sim_rank_fused, sim_rank_topk, tensor_core_supported  # noqa: F401
from .loss_more import BCELossHead, multilabel_scores  # noqa: F401
from .loss_head import (  # noqa: F401
    LOSS_HEADS_REGISTRY,
    CELossHead,
    ClassificationHead,
    LossHead,
    VACELossHead,
    VALCELossHead,
    build_loss_head,
    install_into_reference,
)

__version__ = "0.1.0"
