"""Host-side mirror of gr.iti.mklab.visual.dimreduction.PCA's projection path: the step between
the VLAD vector and the index in the authors' pipeline (ImageVectorization.java:200-207).

  PCA.loadPCAFromFile       J/dimreduction/PCA.java:257-318   text file: line 1 means, line 2 eigenvalues, then one
                                                              eigenvector per line, space-separated
  PCA.sampleToEigenSpace    J/dimreduction/PCA.java:188-208   y = V_t (x - mean); with whitening V_t is pre-multiplied
                                                              by diag(eigenvalue^-0.5) and y is L2-normalised

Learning the basis (addSample / computeBasis, EJML SVD) is out of scope.  The projection runs on the device
(mmidx_pca_project: a tiled binary64 product, the products of a row added for j ascending -- the row-times-vector loop);
EJML is not vendored, so whether that is its exact summation order is unverified and parity with the Java path is claimed
within the 1e-4 relative tolerance only; the final L2 step is the bit-identical one of Normalization.normalizeL2.
The file format and the whitening fold-in are host logic, as in the reference."""
import ctypes as C

import numpy as np

from ._capi import check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PCA:
    def __init__(self, numComponents, numSamples, sampleSize, doWhitening=False, device=-1):
        self.device = device
        self.numComponents, self.numSamples, self.sampleSize = int(numComponents), int(numSamples), int(sampleSize)
        self.doWhitening = bool(doWhitening)
        self.means = None
        self.V_t = None
        self.isPcaInitialized = False

    def loadPCAFromFile(self, source):
        """`source`: path of a file written by PCA.savePCAToFile, or (means[sampleSize], eigenvalues[>= numComponents],
        V_t[>= numComponents][sampleSize]) arrays."""
        if isinstance(source, str):
            with open(source) as f:
                means = np.array(f.readline().strip().split(" "), dtype=np.float64)
                eig = np.array(f.readline().strip().split(" "), dtype=np.float64)
                rows = []
                for _ in range(self.numComponents):
                    line = f.readline()
                    if not line:
                        raise ValueError("Check whether the given PCA matrix contains the correct number of components!")
                    rows.append(np.array(line.strip().split(" ")[: self.sampleSize], dtype=np.float64))
                V = np.stack(rows)
        else:
            means, eig, V = (np.asarray(a, dtype=np.float64) for a in source)
            V = V[: self.numComponents]
        if means.shape != (self.sampleSize,):
            raise ValueError("Means line is wrong!")  # PCA.java:264-266
        if V.shape != (self.numComponents, self.sampleSize):
            raise ValueError("Check whether the given PCA matrix contains the correct number of components!")
        if self.doWhitening:
            if eig.shape[0] < self.numComponents:
                raise ValueError("Eigenvalues line is wrong!")  # PCA.java:277-279
            V = np.power(eig[: self.numComponents], -0.5)[:, None] * V  # W V_t with W = diag(eigenvalue^-0.5)
        self.means, self.V_t = means, np.ascontiguousarray(V)
        self.isPcaInitialized = True

    def sampleToEigenSpaceBatch(self, X):
        if not self.isPcaInitialized:
            raise RuntimeError("PCA is not correctly initiallized!")  # sic, PCA.java:189-192
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self.sampleSize:
            raise ValueError("Unexpected vector length!")
        # y_i = sum_j V_t[i][j] * (x_j - mean_j) on the device; whitening: L2 of every projected row (PCA.java:203-205)
        X = np.ascontiguousarray(X)
        Y = np.empty((X.shape[0], self.numComponents), dtype=np.float64)
        if X.shape[0]:
            check(lib.mmidx_pca_project(_ptr(self.V_t), _ptr(self.means), self.numComponents, self.sampleSize, X.shape[0], _ptr(X),
                                        1 if self.doWhitening else 0, _ptr(Y), self.device))
        return Y

    def sampleToEigenSpace(self, sampleData):
        return self.sampleToEigenSpaceBatch(np.asarray(sampleData, dtype=np.float64).reshape(1, -1))[0]

    def savePCAToFile(self, path, eigenvalues):
        """the file format of PCA.savePCAToFile :219-247 (for fixtures; the un-whitened basis must be passed back in)"""
        if self.isPcaInitialized and self.doWhitening:
            raise RuntimeError("the whitened matrix is not what the file holds")
        with open(path, "w") as f:
            f.write(" ".join(repr(float(x)) for x in self.means) + "\n")
            f.write(" ".join(repr(float(x)) for x in eigenvalues) + "\n")
            for row in self.V_t:
                f.write(" ".join(repr(float(x)) for x in row) + "\n")
