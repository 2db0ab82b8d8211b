"""litepi_b200 -- B200-native backend for the YOLO-LitePi two-stage traffic-sign pipeline.

Drop-in for the reference's ``e2e.py`` wrappers (``NCNNDetector``, ``PyTorchClassifier``,
``HybridPipeline``): same constructor arguments, same per-frame outputs, plus batched
entry points.  All compute is hand-written CUDA for sm_100a behind the C-ABI in
``include/litepi_b200.h``; there is no CPU fallback.
"""
from .ncnn_model import load_ncnn  # noqa: F401
from .detector import B200Detector  # noqa: F401
from .classifier import B200Classifier  # noqa: F401
from .pipeline import B200Pipeline, PipelineMetrics  # noqa: F401
from .evaluate import Evaluator  # noqa: F401

# reference-compatible aliases (src/vntsr/pipeline/e2e.py:195, :350, :399)
NCNNDetector = B200Detector
PyTorchClassifier = B200Classifier
HybridPipeline = B200Pipeline
