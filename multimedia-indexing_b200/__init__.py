"""mmidx-b200: B200 (sm_100a) drop-in for the search / encode hot path of MKLab-ITI/multimedia-indexing.
Importing the package loads libmmidx.so and fails loudly if it is not built (no CPU fallback)."""
from . import _capi
from ._capi import MmidxError, LIB_PATH
from .datastructures import Answer, IVFPQ, PQ, Linear, TransformationType, random_permutation
from .aggregation import VladAggregator, VladAggregatorMultipleVocabularies, normalizeL2, normalizePower, normalizeSSR

__all__ = ["Answer", "IVFPQ", "PQ", "Linear", "TransformationType", "VladAggregator", "VladAggregatorMultipleVocabularies", "normalizeL2", "normalizePower", "normalizeSSR",
           "MmidxError", "random_permutation",
           "LIB_PATH"]


def version():
    return _capi.lib.mmidx_version().decode()
