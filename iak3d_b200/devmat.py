"""Device-resident dense matrices (include/iak3d.h, `iak3d_mat_*`): R's matrix operations for the callers of the hot path
whose matrices are as large as the data blocks (predMatsIAK3D_CLEV_ByBlock, compositeLikIAK3D.R:609-798).  Every
operation runs on the GPU; the host only sequences them."""
import ctypes as C

import numpy as np

from ._lib import NOT_SPD, default_context, dptr, f64


class DeviceMat:
    def __init__(self, rows, cols, ctx=None):
        self.ctx = ctx or default_context()
        self.rows, self.cols = int(rows), int(cols)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.iak3d_mat_create(self.ctx.h, self.rows, self.cols, C.byref(h)))
        self.h = h

    # ---- construction / transfer -------------------------------------------------------------
    @classmethod
    def from_host(cls, a, ctx=None):
        a = f64(np.atleast_2d(a))
        M = cls(a.shape[0], a.shape[1], ctx)
        M.set(a)
        return M

    def set(self, a, r0=0, c0=0):
        a = f64(np.atleast_2d(a))
        self.ctx.check(self.ctx.lib.iak3d_mat_set(self.h, r0, c0, a.shape[0], a.shape[1], dptr(a), a.shape[0]))
        return self

    def to_host(self, r0=0, c0=0, nr=None, nc=None):
        nr = self.rows - r0 if nr is None else nr
        nc = self.cols - c0 if nc is None else nc
        out = np.empty((nr, nc), order="F")
        self.ctx.check(self.ctx.lib.iak3d_mat_get(self.h, r0, c0, nr, nc, dptr(out), max(nr, 1)))
        return out

    # ---- operations ----------------------------------------------------------------------------
    def assign(self, src, dr0=0, dc0=0, sr0=0, sc0=0, nr=None, nc=None, transpose=False, alpha=1.0, beta=0.0):
        """self[dr0:dr0+nr, dc0:dc0+nc] = beta * self[...] + alpha * src[sr0:.., sc0:..] (or its transpose)."""
        if nr is None:
            nr = (src.cols - sc0) if transpose else (src.rows - sr0)
        if nc is None:
            nc = (src.rows - sr0) if transpose else (src.cols - sc0)
        self.ctx.check(self.ctx.lib.iak3d_mat_copy(self.h, dr0, dc0, src.h, sr0, sc0, nr, nc, 1 if transpose else 0, float(alpha),
                                                   float(beta)))
        return self

    def sub(self, r0, c0, nr, nc, transpose=False, alpha=1.0):
        """New matrix = alpha * self[r0:r0+nr, c0:c0+nc] (or its transpose)."""
        out = DeviceMat(nc, nr, self.ctx) if transpose else DeviceMat(nr, nc, self.ctx)
        if transpose:
            out.assign(self, sr0=r0, sc0=c0, nr=nc, nc=nr, transpose=True, alpha=alpha)
        else:
            out.assign(self, sr0=r0, sc0=c0, nr=nr, nc=nc, alpha=alpha)
        return out

    def gather_blocks(self, row_ranges, col_ranges):
        """self[c(rows...), c(cols...)] for lists of contiguous (start, length) ranges -> new matrix."""
        nr = sum(l for _, l in row_ranges); nc = sum(l for _, l in col_ranges)
        out = DeviceMat(nr, nc, self.ctx)
        ro = 0
        for r0, rl in row_ranges:
            co = 0
            for c0, cl in col_ranges:
                if rl and cl:
                    out.assign(self, dr0=ro, dc0=co, sr0=r0, sc0=c0, nr=rl, nc=cl)
                co += cl
            ro += rl
        return out

    def t(self):
        return self.sub(0, 0, self.rows, self.cols, transpose=True)

    def symmetrised(self):
        """0.5 * (A + t(A))."""
        out = DeviceMat(self.rows, self.cols, self.ctx)
        out.assign(self, alpha=0.5)
        out.assign(self, transpose=True, alpha=0.5, beta=1.0)
        return out

    def gemm_nt(self, A, B, alpha=1.0, accumulate=False):
        """self = [self +] alpha * A %*% t(B)."""
        self.ctx.check(self.ctx.lib.iak3d_mat_gemm_nt(self.h, float(alpha), A.h, B.h, 1 if accumulate else 0))
        return self

    @staticmethod
    def matmul_nt(A, B, alpha=1.0):
        out = DeviceMat(A.rows, B.rows, A.ctx)
        return out.gemm_nt(A, B, alpha)

    def spd_inverse(self):
        """In place chol2inv(chol(self)); raises like R's chol() for a matrix that is not positive definite."""
        st = C.c_int32(0)
        self.ctx.check(self.ctx.lib.iak3d_mat_spd_inverse(self.h, C.byref(st)), allow=(NOT_SPD,))
        if st.value != 0:
            raise np.linalg.LinAlgError("the leading minor is not positive definite")
        return self

    def coldot(self, other):
        out = np.empty(self.cols)
        self.ctx.check(self.ctx.lib.iak3d_mat_coldot(self.h, other.h, dptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.lib.iak3d_mat_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cov_mat(setupMats, P):
    """setCIAK3D(...)$C of a resident dataset as a DeviceMat (iak3d_mat_cov_build); None for invalid parameters."""
    from ._lib import BAD_PARS
    M = DeviceMat(setupMats.n, setupMats.n, setupMats.ctx)
    st = M.ctx.check(M.ctx.lib.iak3d_mat_cov_build(setupMats.h, C.byref(P), M.h), allow=(BAD_PARS,))
    if st == BAD_PARS:
        M.close()
        return None
    return M
