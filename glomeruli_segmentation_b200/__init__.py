"""espnet-b200: the ESPNet glomerular-segmentation inference hot path of
jinseikenai/glomeruli_segmentation as hand-written sm_100a CUDA kernels behind a C ABI
(include/espnet_b200.h), with drop-in torch modules (`Model.ESPNet`, `Model.ESPNet_Encoder`),
the tile -> slide stitcher (`wsi`) and the confusion-matrix IoU (`IOUEval.iouEval`)."""
from . import _lib  # noqa: F401
from .Model import ESPNet, ESPNet_Encoder, ESPNetEnsemble, GraphedSegmenter, HostPipeline, FOLD_MEAN_STD  # noqa: F401
from .IOUEval import iouEval  # noqa: F401
from . import wsi  # noqa: F401
from . import frontend  # noqa: F401

__version__ = "0.1.0"
