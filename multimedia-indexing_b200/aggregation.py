"""Host-side mirror of gr.iti.mklab.visual.aggregation.VladAggregator over libmmidx's C ABI.

  AbstractFeatureAggregator.aggregate(double[][])            J/aggregation/AbstractFeatureAggregator.java:72-79
  VladAggregator.aggregateInternal                           J/aggregation/VladAggregator.java:56-70
  AbstractFeatureAggregator.computeNearestCentroid           AFA.java:136-155
`aggregateBatch` is the batched form (many images per launch); `aggregate` is a batch of one."""
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
