"""Host-side mirror of gr.iti.mklab.visual.dimreduction.PCA's projection path: the step between
the VLAD vector and the index in the authors' pipeline (ImageVectorization.java:200-207).

  PCA.loadPCAFromFile       J/dimreduction/PCA.java:257-318   text file: line 1 means, line 2 eigenvalues, then one
                                                              eigenvector per line, space-separated
  PCA.sampleToEigenSpace    J/dimreduction/PCA.java:188-208   y = V_t (x - mean); with whitening V_t is pre-multiplied
                                                              by diag(eigenvalue^-0.5) and y is L2-normalised

Learning the basis (addSample / computeBasis, EJML SVD) is out of scope.  The projection is evaluated as the plain
row-times-vector loop (products added for j ascending); EJML is not vendored, so whether that is its exact summation
order is unverified -- parity is claimed within the 1e-4 relative tolerance only; the final L2 step is the bit-identical
one of aggregation.normalizeL2.  Plain numpy: the reference runs this on the CPU as well."""
import numpy as np

from .aggregation import normalizeL2


class PCA:
    def __init__(self, numComponents, numSamples, sampleSize, doWhitening=False):
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
        # y_i = sum_j V_t[i][j] * (x_j - mean_j), products rounded then added for j ascending (the plain row-times-vector
        # loop of EJML's MatrixVectorMult; np.cumsum adds sequentially) -- deterministic, and the same for one vector
        # or a batch
        Xc = X - self.means[None, :]
        Y = np.empty((X.shape[0], self.numComponents), dtype=np.float64)
        step = max(1, (1 << 23) // max(1, self.numComponents * self.sampleSize))
        for b in range(0, X.shape[0], step):
            prod = Xc[b:b + step, None, :] * self.V_t[None, :, :]
            Y[b:b + step] = np.cumsum(prod, axis=2)[:, :, -1]
        return normalizeL2(Y) if self.doWhitening else Y

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
