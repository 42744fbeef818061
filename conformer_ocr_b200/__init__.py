"""conformer_ocr_b200 — B200-native (sm_100a) implementation of conformer_ocr's inference hot path.

The host side mirrors the reference's interfaces for this path
(`conformer_ocr.conformer.encoder.ConformerEncoder`, `conformer_ocr.pred.PytorchRecognitionModel`,
`conformer_ocr.model.RecognitionModel.forward`) and calls hand-written CUDA kernels through the
C ABI declared in include/cocr_b200.h (libcocr_b200.so, built by `__graft_entry__.build()`).
There is no CPU fallback: without the library or without a Blackwell GPU every compute call raises.
"""
from .conformer.encoder import ConformerEncoder  # noqa: F401
from .pred import PytorchRecognitionModel, greedy_decoder  # noqa: F401
from .model import RecognitionModel  # noqa: F401
from .default_specs import RECOGNITION_HYPER_PARAMS  # noqa: F401
from .codec import LabelCodec  # noqa: F401
from .collate import WidthBucketBatchSampler, collate_sequences  # noqa: F401
from .convert import checkpoint_to_safetensors  # noqa: F401

__all__ = ['ConformerEncoder', 'PytorchRecognitionModel', 'RecognitionModel', 'greedy_decoder', 'RECOGNITION_HYPER_PARAMS',
           'LabelCodec', 'WidthBucketBatchSampler', 'collate_sequences', 'checkpoint_to_safetensors']
