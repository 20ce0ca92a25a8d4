"""Drop-in replacement for the reference's ``model.py``: ``from model import VQABaselineNet,
HierarchicalCoAttentionNet`` (reference main.py:15) resolves to the B200-native modules."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("visual-question-answering_b200")

HierarchicalCoAttentionNet = _pkg.HierarchicalCoAttentionNet
VQABaselineNet = _pkg.VQABaselineNet
QuestionCoAttentionEncoder = _pkg.QuestionCoAttentionEncoder
PhraseConvPool = _pkg.PhraseConvPool
ParallelCoAttention = _pkg.ParallelCoAttention
MLPClassifier = _pkg.MLPClassifier
ImageCoAttentionEncoder = _pkg.ImageCoAttentionEncoder
ImageBaselineEncoder = _pkg.ImageBaselineEncoder
QuestionBaselineEncoder = _pkg.QuestionBaselineEncoder
