"""B200-native AVFormer transformer hot path (drop-in for the reference's models/avformer.py stack).

Public surface mirrors the reference's model files:
    TwoStreamAuralVisualFormer, AudioFormer, VisualFormer          (models/avformer.py)
    VideoModel, ResFormer, TFormer, BasicBlock                     (models/vformer.py)
    AU_former, former_AU_head / tformer_AU_head                    (models/heads.py, models/tformer.py)
    VA_former, SFormerBlock, TFormer(dim=1536)                     (the other variants' instantiations of the same block, inference only)
    Transformer, Attention, FeedForward, PreNorm, Residual, GELU   (models/heads.py:164-256)
    AULoss                                                         (models/loss.py:63-103)
plus ``functional`` (tensor-level wrappers of the C ABI), ``autograd`` (the backward bridges), ``FusedAdam``
(train.py:334's optimiser as one kernel over a flat bucket), ``dp`` (data-parallel sharding helpers) and ``build()``.
"""
from . import _lib
from . import functional
from . import autograd
from . import optim
from . import graphs
from . import dp
from ._lib import build
from .audio import AudioModel
from .avformer import AudioFormer, TwoStreamAuralVisualFormer, VisualFormer, load_pretrain
from .encoder import GELU, Attention, FeedForward, PreNorm, Residual, Transformer, default_precision, set_default_precision
from .heads import AU_former, VA_former, former_AU_head, tformer_AU_head
from .loss import AULoss
from .metrics import MultiLabelAccF1
from . import outputs
from .outputs import AUResultWriter
from .optim import FusedAdam
from .graphs import GraphedHotPath, GraphedTrainStep
from .inference import InferenceEngine
from .video import BasicBlock, Dummy, ResFormer, SFormerBlock, TFormer, VideoModel

__all__ = [
    "TwoStreamAuralVisualFormer", "AudioFormer", "VisualFormer", "VideoModel", "ResFormer", "TFormer", "BasicBlock", "Dummy",
    "AU_former", "VA_former", "SFormerBlock", "former_AU_head", "tformer_AU_head", "Transformer", "Attention", "FeedForward", "PreNorm", "Residual", "GELU",
    "AULoss", "MultiLabelAccF1", "AUResultWriter", "outputs", "AudioModel", "FusedAdam", "GraphedHotPath", "GraphedTrainStep", "InferenceEngine", "graphs", "dp", "functional", "autograd", "optim", "build", "set_default_precision", "default_precision", "load_pretrain",
]
