"""TEST INFRASTRUCTURE ONLY -- stand-in for `petsc4py.PETSc` (see package doc)."""
import numpy as np

COMM_SELF = 'PETSC_COMM_SELF'


class Vec:
    def createWithArray(self, array, comm=None):
        self.array = array  # aliases the caller's array, like PETSc does
        return self

    def setArray(self, array):
        self.array[...] = array


class Mat:
    class SORType:
        FORWARD_SWEEP = 1
        BACKWARD_SWEEP = 2

    def createAIJWithArrays(self, size, csr, comm=None):
        indptr, indices, data = csr
        self.n = size[0]
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        return self

    def SOR(self, b, x, its=1, sortype=None):
        from oracle import cgs
        assert x.array.flags.c_contiguous and x.array.dtype == np.float64
        cgs.gauss_seidel(self.indptr, self.indices, self.data, b.array,
                         x.array, its,
                         backward=(sortype == Mat.SORType.BACKWARD_SWEEP))
