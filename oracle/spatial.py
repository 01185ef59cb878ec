"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the spatial side of the reference's hot path:
  MatrixFreeOperator  include/operators.h:967-1191  (cell loop :1112-1133, cell kernel :1135-1173,
                      diagonal :1092-1110, assembled matrix :1020-1033, coefficient :1060-1087)
  SystemMatrix        include/operators.h:536-640   (vmult, Tvmult, vmult_slice_add, diagonal)
  Coefficient         include/operators.h:870-965
  mesh / FE / quadrature set-up as used by the driver tests/tp_01.cc:77-100
deal.II semantics restated from SURVEY.md App. A (FE_Q on Gauss-Lobatto nodes, QGauss(k+1),
MappingQ1, homogeneous Dirichlet constraints: constrained DoFs read as 0 and are never written).

Numbering: lexicographic global DoF grid (x fastest) on a structured box mesh of n[0] x n[1] (x n[2])
cells — the numbering the product uses.  deal.II's own numbering differs by a permutation only.
"""
import numpy as np
import scipy.sparse as sp

from . import quadrature as Q


class Mesh:
    """subdivided_hyper_rectangle(subdivisions, p1, p2) + refine_global(r) (tests/tp_01.cc:83-88);
    optional vertex perturbation stands in for GridTools::distort_random (tp_01.cc:89-90; the deal.II
    random stream is not reproducible without deal.II, SURVEY.md A.10 — both sides of the parity
    tests use THIS generator)."""

    def __init__(self, dim, subdivisions, refinement, lower=None, upper=None, distort=0.0, seed=1234):
        self.dim = dim
        self.subdivisions = list(subdivisions)
        self.n = [int(s) * (1 << refinement) for s in subdivisions]
        self.lower = np.zeros(dim) if lower is None else np.asarray(lower, float)
        self.upper = np.ones(dim) if upper is None else np.asarray(upper, float)
        axes = [np.linspace(self.lower[d], self.upper[d], self.n[d] + 1) for d in range(dim)]
        # vertices[(iz,) iy, ix, comp]
        grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
        self.vertices = np.stack(grids, axis=-1).astype(float)
        self.cartesian = distort == 0.0
        if distort != 0.0:
            h = np.array([(self.upper[d] - self.lower[d]) / self.n[d] for d in range(dim)])
            rng = np.random.RandomState(seed)
            shift = rng.uniform(-1.0, 1.0, size=self.vertices.shape) * distort * h.min()
            interior = np.ones(self.vertices.shape[:-1], bool)
            for d in range(dim):
                ax = dim - 1 - d
                sl = [slice(None)] * dim
                sl[ax] = 0
                interior[tuple(sl)] = False
                sl[ax] = -1
                interior[tuple(sl)] = False
            self.vertices = self.vertices + shift * interior[..., None]
        self.spc_step = min((self.upper[d] - self.lower[d]) / self.subdivisions[d] for d in range(dim))

    @property
    def n_cells(self):
        return int(np.prod(self.n))

    def coarsen(self):
        """One level of the geometric coarsening sequence (tests/tp_01.cc:171-174): every second vertex."""
        m = Mesh.__new__(Mesh)
        m.dim, m.subdivisions = self.dim, self.subdivisions
        assert all(v % 2 == 0 for v in self.n)
        m.n = [v // 2 for v in self.n]
        m.lower, m.upper, m.cartesian, m.spc_step = self.lower, self.upper, self.cartesian, self.spc_step
        sl = tuple([slice(None, None, 2)] * self.dim)
        m.vertices = self.vertices[sl].copy()
        return m

    def cell_vertices(self):
        """[n_cells, 2^dim, dim] in deal.II (lexicographic) vertex order, cells lexicographic."""
        d = self.dim
        out = []
        for v in range(1 << d):
            sl = []
            for ax in range(d):            # array axis ax <-> coordinate d-1-ax
                bit = (v >> (d - 1 - ax)) & 1
                sl.append(slice(bit, bit + self.n[d - 1 - ax]))
            out.append(self.vertices[tuple(sl)].reshape(-1, d))
        return np.stack(out, axis=1)


class Space:
    """FE_Q(k) DoFs on a Mesh with zero Dirichlet boundary (tests/tp_01.cc:77,92-100)."""

    def __init__(self, mesh, degree, dirichlet_faces=None):
        """dirichlet_faces: bit 2*d+side set = zero Dirichlet on that face (default: all faces)."""
        self.mesh, self.k, self.dim = mesh, degree, mesh.dim
        if dirichlet_faces is None:
            dirichlet_faces = (1 << (2 * mesh.dim)) - 1
        self.n1 = degree + 1
        self.np = [degree * v + 1 for v in mesh.n]
        self.n_dofs = int(np.prod(self.np))
        self.gll = Q.gauss_lobatto(self.n1)[0]
        self.xq, self.wq = Q.gauss(self.n1)                       # QGauss(k+1), tp_01.cc:78
        self.S = Q.lagrange_eval(self.gll, self.xq).T.copy()      # S[q,i] = phi_i(x_q)
        self.D = Q.lagrange_deriv(self.gll, self.xq).T.copy()     # D[q,i] = phi_i'(x_q)
        self._dirichlet_faces = dirichlet_faces
        self._constrained = None
        self._cell_dofs = None
        self._geom = None

    # the index tables are built on first use: the CPU baseline of bench.py runs 96^3-cell bricks through
    # oracle/cpu_ref.cpp, which needs neither of them
    @property
    def constrained(self):
        if self._constrained is None:
            d = self.dim
            mask = np.zeros(self.np[::-1], bool)
            for ax in range(d):
                coord = d - 1 - ax                                  # array axis ax <-> coordinate direction
                sl = [slice(None)] * d
                if (self._dirichlet_faces >> (2 * coord)) & 1:
                    sl[ax] = 0
                    mask[tuple(sl)] = True
                if (self._dirichlet_faces >> (2 * coord + 1)) & 1:
                    sl[ax] = -1
                    mask[tuple(sl)] = True
            self._constrained = mask.reshape(-1)
        return self._constrained

    @property
    def cell_dofs(self):
        """cell_dofs[c, local], lexicographic in both"""
        if self._cell_dofs is None:
            mesh, d, k = self.mesh, self.dim, self.k
            loc = np.arange(self.n1)
            if d == 2:
                cy, cx = np.meshgrid(np.arange(mesh.n[1]), np.arange(mesh.n[0]), indexing="ij")
                iy = (k * cy.reshape(-1))[:, None, None] + loc[None, :, None]
                ix = (k * cx.reshape(-1))[:, None, None] + loc[None, None, :]
                self._cell_dofs = (ix + self.np[0] * iy).reshape(mesh.n_cells, -1)
            else:
                cz, cy, cx = np.meshgrid(np.arange(mesh.n[2]), np.arange(mesh.n[1]), np.arange(mesh.n[0]),
                                         indexing="ij")
                iz = (k * cz.reshape(-1))[:, None, None, None] + loc[None, :, None, None]
                iy = (k * cy.reshape(-1))[:, None, None, None] + loc[None, None, :, None]
                ix = (k * cx.reshape(-1))[:, None, None, None] + loc[None, None, None, :]
                self._cell_dofs = (ix + self.np[0] * (iy + self.np[1] * iz)).reshape(mesh.n_cells, -1)
        return self._cell_dofs

    # ---- geometry (MappingQ1)
    def geometry(self):
        """Per cell and q-point: Ginv[c,q,a,b] = (J^-1 J^-T)_{ab} * JxW and JxW[c,q]; q lexicographic."""
        if self._geom is not None:
            return self._geom
        d = self.dim
        X = self.mesh.cell_vertices()                                     # [C, 2^d, d]
        xq = self.xq
        # 1D linear shape values / derivatives at q points
        N1 = np.stack([1.0 - xq, xq], axis=0)                              # [2, nq]
        dN1 = np.stack([-np.ones_like(xq), np.ones_like(xq)], axis=0)
        nq = len(xq)
        if d == 2:
            # vertex v = vx + 2*vy ; q = qx + nq*qy
            Nv = np.zeros((4, nq * nq))
            dNv = np.zeros((4, 2, nq * nq))
            for vy in range(2):
                for vx in range(2):
                    v = vx + 2 * vy
                    Nv[v] = np.einsum("y,x->yx", N1[vy], N1[vx]).reshape(-1)
                    dNv[v, 0] = np.einsum("y,x->yx", N1[vy], dN1[vx]).reshape(-1)
                    dNv[v, 1] = np.einsum("y,x->yx", dN1[vy], N1[vx]).reshape(-1)
            w = np.einsum("y,x->yx", self.wq, self.wq).reshape(-1)
        else:
            Nv = np.zeros((8, nq ** 3))
            dNv = np.zeros((8, 3, nq ** 3))
            for vz in range(2):
                for vy in range(2):
                    for vx in range(2):
                        v = vx + 2 * vy + 4 * vz
                        Nv[v] = np.einsum("z,y,x->zyx", N1[vz], N1[vy], N1[vx]).reshape(-1)
                        dNv[v, 0] = np.einsum("z,y,x->zyx", N1[vz], N1[vy], dN1[vx]).reshape(-1)
                        dNv[v, 1] = np.einsum("z,y,x->zyx", N1[vz], dN1[vy], N1[vx]).reshape(-1)
                        dNv[v, 2] = np.einsum("z,y,x->zyx", dN1[vz], N1[vy], N1[vx]).reshape(-1)
            w = np.einsum("z,y,x->zyx", self.wq, self.wq, self.wq).reshape(-1)
        J = np.einsum("cva,vbq->cqab", X, dNv)                             # dx_a/dxi_b
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)                                            # dxi_a/dx_b
        JxW = det * w[None, :]
        G = np.einsum("cqab,cqdb->cqad", Jinv, Jinv) * JxW[..., None, None]
        points = np.einsum("cva,vq->cqa", X, Nv)
        self._geom = (G, JxW, points)
        return self._geom


def _interp(S, u, dim):
    """apply S (q x i) along every tensor direction of u[..., (z,) y, x]."""
    if dim == 2:
        return np.einsum("ay,bx,...yx->...ab", S, S, u, optimize=True)
    return np.einsum("az,by,cx,...zyx->...abc", S, S, S, u, optimize=True)


class MatrixFreeOperator:
    """mass_scaling*M + laplace_scaling*K, optional per-(cell,q) coefficients
    (include/operators.h:967-1191)."""

    def __init__(self, space, mass_scaling, laplace_scaling, dtype=np.float64):
        self.sp, self.ms, self.ls, self.dtype = space, mass_scaling, laplace_scaling, dtype
        self.mass_coeff = None
        self.laplace_coeff = None
        self._diag = None

    def evaluate_coefficient(self, fun):
        """operators.h:1060-1087: coefficient sampled at the q-points of every cell."""
        _, _, pts = self.sp.geometry()
        val = fun(pts)
        if self.ms != 0.0:
            self.mass_coeff = val
        if self.ls != 0.0:
            self.laplace_coeff = val

    def m(self):
        return self.sp.n_dofs

    def cell_apply(self, ucell):
        """do_cell_integral_local (operators.h:1135-1173) on ucell[C, n_c] -> [C, n_c]."""
        s = self.sp
        d, n1 = s.dim, s.n1
        dt = self.dtype
        G, JxW, _ = s.geometry()
        S, D = s.S.astype(dt), s.D.astype(dt)
        C = ucell.shape[0]
        u = ucell.reshape((C,) + (n1,) * d).astype(dt)
        out = np.zeros_like(u)
        if self.ms != 0.0:
            uq = _interp(S, u, d).reshape(C, -1)
            c = self.mass_coeff if self.mass_coeff is not None else self.ms
            vq = (uq * (c * JxW).astype(dt)).reshape((C,) + (n1,) * d)
            out += _interp(S.T.copy(), vq, d)
        if self.ls != 0.0:
            c = self.laplace_coeff if self.laplace_coeff is not None else self.ls
            Gc = (G * (np.asarray(c)[..., None, None] if np.ndim(c) else c)).astype(dt)
            if d == 2:
                gx = np.einsum("ay,bx,cyx->cab", S, D, u, optimize=True).reshape(C, -1)
                gy = np.einsum("ay,bx,cyx->cab", D, S, u, optimize=True).reshape(C, -1)
                g = np.stack([gx, gy], axis=-1)
                t = np.einsum("cqab,cqb->cqa", Gc, g).reshape((C, n1, n1, 2))
                out += np.einsum("ay,bx,cab->cyx", S, D, t[..., 0], optimize=True)
                out += np.einsum("ay,bx,cab->cyx", D, S, t[..., 1], optimize=True)
            else:
                gx = np.einsum("az,by,ex,czyx->cabe", S, S, D, u, optimize=True).reshape(C, -1)
                gy = np.einsum("az,by,ex,czyx->cabe", S, D, S, u, optimize=True).reshape(C, -1)
                gz = np.einsum("az,by,ex,czyx->cabe", D, S, S, u, optimize=True).reshape(C, -1)
                g = np.stack([gx, gy, gz], axis=-1)
                t = np.einsum("cqab,cqb->cqa", Gc, g).reshape((C, n1, n1, n1, 3))
                out += np.einsum("az,by,ex,cabe->czyx", S, S, D, t[..., 0], optimize=True)
                out += np.einsum("az,by,ex,cabe->czyx", S, D, S, t[..., 1], optimize=True)
                out += np.einsum("az,by,ex,cabe->czyx", D, S, S, t[..., 2], optimize=True)
        return out.reshape(C, -1)

    def vmult(self, src):
        """cell_loop(do_cell_integral_range, zero_dst=true): operators.h:1013-1018, 1112-1133.
        read_dof_values: constrained -> 0; distribute_local_to_global: constrained rows skipped."""
        s = self.sp
        srcm = np.where(s.constrained, 0, src).astype(self.dtype)
        ucell = srcm[s.cell_dofs]
        vcell = self.cell_apply(ucell)
        dst = np.zeros(s.n_dofs, dtype=self.dtype)
        np.add.at(dst, s.cell_dofs.reshape(-1), vcell.reshape(-1))
        dst[s.constrained] = 0
        return dst

    def cell_matrices(self):
        """dense cell matrices [C, n_c, n_c] (columns = images of unit vectors; no constraints)."""
        s = self.sp
        nc = s.cell_dofs.shape[1]
        C = s.cell_dofs.shape[0]
        out = np.zeros((C, nc, nc), dtype=self.dtype)
        for j in range(nc):
            e = np.zeros((C, nc), dtype=self.dtype)
            e[:, j] = 1
            out[:, :, j] = self.cell_apply(e)
        return out

    def compute_system_matrix(self):
        """MatrixFreeTools::compute_matrix with zero-Dirichlet AffineConstraints
        (operators.h:1020-1033): constrained rows/cols removed, the diagonal of a
        constrained row receives the (positive) sum of the local diagonal entries."""
        s = self.sp
        Ac = self.cell_matrices()
        C, nc, _ = Ac.shape
        rows = np.repeat(s.cell_dofs[:, :, None], nc, axis=2)
        cols = np.repeat(s.cell_dofs[:, None, :], nc, axis=1)
        cr = s.constrained[rows]
        cc = s.constrained[cols]
        vals = np.where(cr | cc, 0.0, Ac)
        diag_fix = cr & (rows == cols)
        vals = np.where(diag_fix, np.abs(Ac), vals)
        A = sp.coo_matrix((vals.reshape(-1), (rows.reshape(-1), cols.reshape(-1))),
                          shape=(s.n_dofs, s.n_dofs)).tocsr()
        A.sum_duplicates()
        return A

    def compute_diagonal(self):
        """operators.h:1092-1110."""
        if self._diag is None:
            s = self.sp
            Ac = self.cell_matrices()
            dloc = np.einsum("cii->ci", Ac)
            d = np.zeros(s.n_dofs, dtype=self.dtype)
            np.add.at(d, s.cell_dofs.reshape(-1), dloc.reshape(-1))
            d[s.constrained] = 0
            tol = np.sqrt(np.finfo(self.dtype).eps)
            dinv = np.where(np.abs(d) > tol, 1.0 / np.where(d == 0, 1, d), 1.0).astype(self.dtype)
            self._diag = (d, dinv)
        return self._diag


class SystemMatrix:
    """include/operators.h:516-663.  Block vectors are arrays [nb, N]."""

    def __init__(self, K, M, Alpha, Beta):
        self.K, self.M = K, M
        self.Alpha = np.atleast_2d(np.asarray(Alpha))
        self.Beta = np.atleast_2d(np.asarray(Beta))
        self.alpha_is_zero = not np.any(self.Alpha)
        self.beta_is_zero = not np.any(self.Beta)
        self.dtype = K.dtype

    def n_blocks(self):
        return self.Alpha.shape[0]

    def vmult(self, src):
        """operators.h:536-559 (the unfused reference algorithm: 2*nb cell loops + axpys)."""
        nb = src.shape[0]
        assert self.Alpha.shape[0] == nb
        dst = np.zeros_like(src)
        for i in range(nb):
            tmp = self.K.vmult(src[i])
            for j in range(nb):
                if self.Alpha[j, i] != 0.0:
                    dst[j] += self.dtype(self.Alpha[j, i]) * tmp
            tmp = self.M.vmult(src[i])
            for j in range(nb):
                if self.Beta[j, i] != 0.0:
                    dst[j] += self.dtype(self.Beta[j, i]) * tmp
        return dst

    def Tvmult(self, src):
        """operators.h:561-583."""
        nb = src.shape[0]
        dst = np.zeros_like(src)
        for i in range(nb):
            tmp = self.K.vmult(src[i])
            for j in range(nb):
                if self.Alpha[i, j] != 0.0:
                    dst[j] += self.dtype(self.Alpha[i, j]) * tmp
            tmp = self.M.vmult(src[i])
            for j in range(nb):
                if self.Beta[i, j] != 0.0:
                    dst[j] += self.dtype(self.Beta[i, j]) * tmp
        return dst

    def vmult_slice_add(self, dst, src0):
        """operators.h:586-611 (nb x 1 matrices; src0 is ONE spatial vector)."""
        nb = dst.shape[0]
        if not self.alpha_is_zero:
            tmp = self.K.vmult(src0)
            for j in range(nb):
                if self.Alpha[j, 0] != 0.0:
                    dst[j] += self.dtype(self.Alpha[j, 0]) * tmp
        if not self.beta_is_zero:
            tmp = self.M.vmult(src0)
            for j in range(nb):
                if self.Beta[j, 0] != 0.0:
                    dst[j] += self.dtype(self.Beta[j, 0]) * tmp
        return dst

    def vmult_slice(self, src0, nb=None):
        """operators.h:377-382."""
        nb = self.Alpha.shape[0] if nb is None else nb
        dst = np.zeros((nb, self.K.m()), dtype=self.dtype)
        return self.vmult_slice_add(dst, src0)

    def get_matrix_diagonal(self):
        """operators.h:613-625: diag_i = Alpha(i,i) diag K + Beta(i,i) diag M."""
        dK = self.K.compute_diagonal()[0]
        dM = self.M.compute_diagonal()[0]
        return np.stack([self.Alpha[i, i] * dK + self.Beta[i, i] * dM for i in range(self.Alpha.shape[0])])


class Coefficient:
    """include/operators.h:870-965: c1 (y<0.2), c2 (y>=0.2, x<0.2), c3 otherwise, times a
    per-coarse-cell factor U(1-dc, 1+dc).  The reference draws the factors from boost::mt19937
    (default seed 5489) through boost::uniform_real_distribution, i.e. ONE 32-bit draw per value:
    x = a + (b-a) * u / 2^32, filled in Table order (last index fastest)."""

    def __init__(self, dim, subdivisions, lower, upper, distort_coeff=0.0, c1=1.0, c2=9.0, c3=16.0):
        self.dim, self.c = dim, (c1, c2, c3)
        self.lower = np.asarray(lower, float)
        self.distorted = distort_coeff != 0.0
        if self.distorted:
            n = int(np.prod(subdivisions))
            rs = np.random.RandomState(5489)        # MT19937 with the standard init_genrand(5489)
            u = rs._bit_generator.random_raw(n).astype(np.float64)   # raw genrand_int32 stream
            a, b = 1 - distort_coeff, 1 + distort_coeff
            vals = a + (b - a) * (u / 4294967296.0)
            self.table = vals.reshape(list(subdivisions))      # Table(i0, i1(, i2)), last fastest
            self.step = (np.asarray(upper, float) - self.lower) / np.asarray(subdivisions, float)

    def __call__(self, pts):
        px, py = pts[..., 0], pts[..., 1]
        c1, c2, c3 = self.c
        v = np.where(py >= 0.2, np.where(px < 0.2, c2, c3), c1).astype(float)
        if self.distorted:
            idx = tuple(((pts[..., a] - self.lower[a]) / self.step[a]).astype(np.int64).clip(
                0, self.table.shape[a] - 1) for a in range(self.dim))
            v = v * self.table[idx]
        return v
