"""Host-side mirror of gr.iti.mklab.visual.aggregation.VladAggregator over libmmidx's C ABI.

  AbstractFeatureAggregator.aggregate(double[][])            J/aggregation/AbstractFeatureAggregator.java:72-79
  VladAggregator.aggregateInternal                           J/aggregation/VladAggregator.java:56-70
  AbstractFeatureAggregator.computeNearestCentroid           AFA.java:136-155
`aggregateBatch` is the batched form (many images per launch); `aggregate` is a batch of one.

Row f3, the step right after VLAD accumulation, stays on the host exactly as in the reference:
  Normalization.normalizeL2 / normalizePower / normalizeSSR  J/utilities/Normalization.java:21-37, 74-79, 90-94
  VladAggregatorMultipleVocabularies.aggregate               J/aggregation/VladAggregatorMultipleVocabularies.java:84-101
The L2 norm is the reference's index-ascending sum of rounded squares (np.cumsum adds sequentially), so it is
bit-identical; the power step goes through libm pow like Math.pow (both within 1 ulp of the exact value)."""
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


# ---- Normalization.java (host side; vectors [n] or batches [rows][n], returned as new arrays) ----
def normalizeL2(v):
    """Normalization.normalizeL2 :21-37: v / sqrt(sum v_i^2), squares added for i ascending; an all-zero vector is
    filled with 1 (sic)."""
    v = np.array(v, dtype=np.float64, copy=True)
    rows = v.reshape(1, -1) if v.ndim == 1 else v
    sq = rows * rows
    norm = np.sqrt(np.cumsum(sq, axis=1)[:, -1]) if rows.shape[1] else np.zeros(rows.shape[0])
    zero = norm == 0
    rows[~zero] = rows[~zero] / norm[~zero, None]
    rows[zero] = 1.0
    return v


def normalizePower(v, a):
    """Normalization.normalizePower :74-79: signum(x) * pow(|x|, a)"""
    v = np.asarray(v, dtype=np.float64)
    return np.sign(v) * np.power(np.abs(v), a)


def normalizeSSR(v):
    """Normalization.normalizeSSR :90-94: power(0.5) then L2"""
    return normalizeL2(normalizePower(v, 0.5))


class VladAggregatorMultipleVocabularies:
    """J/aggregation/VladAggregatorMultipleVocabularies.java: one VladAggregator per codebook; every sub-VLAD is
    power(0.5)+L2 normalised, the concatenation is L2 normalised again when there is more than one vocabulary."""

    def __init__(self, codebooks, device=-1, aggregator=VladAggregator):
        self.vladAggregators = [aggregator(cb, device) for cb in codebooks]
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
        subs = []
        for agg in self.vladAggregators:
            sub = agg.aggregateBatch(images)[0]
            if self.normalizationsOn:
                sub = normalizeL2(normalizePower(sub, 0.5))
            subs.append(sub)
        multi = np.concatenate(subs, axis=1)
        if len(self.vladAggregators) > 1 and self.normalizationsOn:
            multi = normalizeL2(multi)
        return multi

    def aggregate(self, descriptors):
        descriptors = np.asarray(descriptors, dtype=np.float64)
        return self.aggregateBatch([descriptors])[0]
