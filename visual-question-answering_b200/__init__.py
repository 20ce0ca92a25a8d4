"""hiecoattn_b200 -- B200-native Hierarchical Co-Attention hot path (drop-in for the reference's model.py).

Importing the package never touches the GPU or the shared library; the library is loaded on the first
op call and a missing build raises (there is no CPU fallback).  The directory name carries a hyphen, so
import it with ``importlib.import_module("visual-question-answering_b200")`` or through the root-level
``model.py`` shim, which is what the reference's ``main.py`` does (``from model import ...``).
"""
from . import _lib, dp, modules, ops, optim, staging, synthetic   # noqa: F401
from .modules import (CrossEntropyLoss, HieCoAttnHotPath, InferenceSession, HierarchicalCoAttentionNet, ImageBaselineEncoder, ImageCoAttentionEncoder,  # noqa: F401
                      MLPClassifier, ParallelCoAttention, PhraseConvPool, QuestionBaselineEncoder,
                      QuestionCoAttentionEncoder, QuestionLens, VQABaselineNet)

__all__ = ["HierarchicalCoAttentionNet", "VQABaselineNet", "QuestionCoAttentionEncoder", "PhraseConvPool",
           "ParallelCoAttention", "MLPClassifier", "ImageCoAttentionEncoder", "ImageBaselineEncoder",
           "QuestionBaselineEncoder", "HieCoAttnHotPath", "InferenceSession", "CrossEntropyLoss", "QuestionLens", "ops", "synthetic", "dp", "optim", "staging"]
