"""Host-side mirror of gr.iti.mklab.visual.aggregation.VladAggregator over libmmidx's C ABI.

  AbstractFeatureAggregator.aggregate(double[][])            J/aggregation/AbstractFeatureAggregator.java:72-79
  VladAggregator.aggregateInternal                           J/aggregation/VladAggregator.java:56-70
  AbstractFeatureAggregator.computeNearestCentroid           AFA.java:136-155
`aggregateBatch` is the batched form (many images per launch); `aggregate` is a batch of one.

Row f3, the step right after VLAD accumulation, runs on the device too:
  Normalization.normalizeL2 / normalizePower / normalizeSSR  J/utilities/Normalization.java:21-37, 74-79, 90-94   mmidx_normalize_rows
  VladAggregatorMultipleVocabularies.aggregate               J/aggregation/VladAggregatorMultipleVocabularies.java:84-101   mmidx_vlad_multi
The L2 norm is the reference's index-ascending sum of rounded squares, so it is bit-identical; power(0.5) is the correctly
rounded square root (Math.pow may differ from it in the last bit).  No CPU path here."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import MmidxError, check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class VladAggregator:
    def __init__(self, codebook, device=-1):
        if isinstance(codebook, str):
            # AFA.readQuantizer AFA.java:234-254: lines without a comma are skipped
            codebook = [[float(x) for x in line.strip().split(",")] for line in open(codebook) if "," in line]
        self.codebook = np.ascontiguousarray(codebook, dtype=np.float64)
        self.numCentroids, self.descriptorLength = self.codebook.shape
        self.device = device

    def getVectorLength(self):
        """VladAggregator.getVectorLength: numCentroids * descriptorLength"""
        return self.numCentroids * self.descriptorLength

    def aggregate(self, descriptors):
        descriptors = np.asarray(descriptors, dtype=np.float64)
        if descriptors.size == 0:
            descriptors = descriptors.reshape(0, self.descriptorLength)
        # AFA.java:74-76
        if descriptors.ndim != 2 or descriptors.shape[1] != self.descriptorLength:
            raise MmidxError(_capi.ERR_DIM, "Descriptor length does not match codebook centroid length")
        return self.aggregateBatch([descriptors])[0][0]

    def aggregateBatch(self, images, return_assign=False):
        """images: list of [n_i][D] arrays, or (desc[sum n][D], offsets[n_img+1]). Returns (vlads[n_img][K*D], assign)."""
        if isinstance(images, tuple):
            desc, offsets = images
            desc = np.ascontiguousarray(desc, dtype=np.float64)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        else:
            offsets = np.zeros(len(images) + 1, dtype=np.int64)
            for i, im in enumerate(images):
                offsets[i + 1] = offsets[i] + len(im)
            desc = np.concatenate([np.asarray(im, dtype=np.float64).reshape(-1, self.descriptorLength) for im in images]) \
                if len(images) else np.zeros((0, self.descriptorLength))
            desc = np.ascontiguousarray(desc, dtype=np.float64)
        if desc.ndim != 2 or desc.shape[1] != self.descriptorLength:
            raise MmidxError(_capi.ERR_DIM, "Descriptor length does not match codebook centroid length")
        n_img = offsets.shape[0] - 1
        out = np.empty((n_img, self.getVectorLength()), dtype=np.float64)
        assign = np.empty(desc.shape[0], dtype=np.int32)
        check(lib.mmidx_vlad(_ptr(self.codebook), self.numCentroids, self.descriptorLength, n_img, _ptr(offsets),
                             _ptr(desc), _ptr(out), _ptr(assign), self.device))
        return out, (assign if return_assign else None)


# ---- Normalization.java on the device (mmidx_normalize_rows): vectors [n] or batches [rows][n], returned as new arrays ----
def _normalize(v, do_power, a, do_l2, device=-1):
    v = np.array(v, dtype=np.float64, copy=True, order="C")
    rows = v.reshape(1, -1) if v.ndim == 1 else v.reshape(v.shape[0], -1)
    if rows.shape[1] == 0 or rows.shape[0] == 0:
        return v
    check(lib.mmidx_normalize_rows(_ptr(rows), rows.shape[0], rows.shape[1], 1 if do_power else 0, float(a), 1 if do_l2 else 0, device))
    return v


def normalizeL2(v, device=-1):
    """Normalization.normalizeL2 :21-37: v / sqrt(sum v_i^2), squares added for i ascending (bit-identical norm); an all-zero
    vector is filled with 1 (sic)."""
    return _normalize(v, False, 0.0, True, device)


def normalizePower(v, a, device=-1):
    """Normalization.normalizePower :74-79: signum(x) * pow(|x|, a)"""
    return _normalize(v, True, a, False, device)


def normalizeSSR(v, device=-1):
    """Normalization.normalizeSSR :90-94: power(0.5) then L2, one kernel"""
    return _normalize(v, True, 0.5, True, device)


class VladAggregatorMultipleVocabularies:
    """J/aggregation/VladAggregatorMultipleVocabularies.java: one VladAggregator per codebook over the SAME descriptors; every
    sub-VLAD is power(0.5)+L2 normalised, the concatenation is L2 normalised again when there is more than one vocabulary
    (:84-101).  One device call per batch of images (mmidx_vlad_multi)."""

    def __init__(self, codebooks, device=-1):
        self.vladAggregators = [VladAggregator(cb, device) for cb in codebooks]
        self.device = device
        D = {a.descriptorLength for a in self.vladAggregators}
        if len(D) != 1:
            raise MmidxError(_capi.ERR_DIM, "Descriptor length does not match codebook centroid length")
        self.descriptorLength = D.pop()
        self._stack = np.ascontiguousarray(np.concatenate([a.codebook for a in self.vladAggregators]), dtype=np.float64)
        self._Ks = np.asarray([a.numCentroids for a in self.vladAggregators], dtype=np.int32)
        self.vectorLength = sum(a.getVectorLength() for a in self.vladAggregators)
        self.normalizationsOn = True

    def getVectorLength(self):
        return self.vectorLength

    def isNormalizationsOn(self):
        return self.normalizationsOn

    def setNormalizationsOn(self, on):
        self.normalizationsOn = bool(on)

    def aggregateBatch(self, images):
        """images as for VladAggregator.aggregateBatch; returns multiVLADs [n_img][vectorLength]"""
        if isinstance(images, tuple):
            desc, offsets = images
            desc = np.ascontiguousarray(desc, dtype=np.float64)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        else:
            offsets = np.zeros(len(images) + 1, dtype=np.int64)
            for i, im in enumerate(images):
                offsets[i + 1] = offsets[i] + len(im)
            desc = np.concatenate([np.asarray(im, dtype=np.float64).reshape(-1, self.descriptorLength) for im in images]) \
                if len(images) else np.zeros((0, self.descriptorLength))
            desc = np.ascontiguousarray(desc, dtype=np.float64)
        if desc.ndim != 2 or desc.shape[1] != self.descriptorLength:
            raise MmidxError(_capi.ERR_DIM, "Descriptor length does not match codebook centroid length")
        n_img = offsets.shape[0] - 1
        out = np.empty((n_img, self.vectorLength), dtype=np.float64)
        check(lib.mmidx_vlad_multi(_ptr(self._stack), len(self._Ks), _ptr(self._Ks), self.descriptorLength, n_img, _ptr(offsets),
                                   _ptr(desc), 1 if self.normalizationsOn else 0, _ptr(out), self.device))
        return out

    def aggregate(self, descriptors):
        descriptors = np.asarray(descriptors, dtype=np.float64)
        return self.aggregateBatch([descriptors])[0]
