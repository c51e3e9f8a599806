"""TEST INFRASTRUCTURE: a CPU stand-in for cufinufft_b200.multi.SlabPlan built on the oracle.

Same stage methods (type1_spread / halo_pack / halo_add / type1_finish / type2 / halo_buffers), so
the orchestration code of cufinufft_b200.multi (ring exchange, all-reduce, routing) runs unchanged
on CPU tensors under gloo.  The stages restate csrc/slab.cu in numpy: the separated FFT (z on mode
columns, (x,y) on planes), the haloed local grid, the deferred phihat division.  Spread / interp
/ phihat come from the oracle (reference arithmetic: src/3d/spreadinterp3d.cu, src/common.cu:16-45,
src/deconvolve_wrapper.cu:52-121)."""
import numpy as np

from cufinufft_b200.multi import slab_halo, slab_range
from oracle import oracle as orc


def _mode_to_grid(m, nf):
    k = np.arange(m) - m // 2
    return np.where(k >= 0, k, nf + k)


class OracleSlab:
    def __init__(self, nufft_type, modes, eps, dtype, rank, world, isign=None):
        self.type, self.rank, self.world = nufft_type, rank, world
        self.dtype = np.dtype(dtype)
        self.cd = np.complex64 if self.dtype == np.float32 else np.complex128
        self.modes = tuple(modes)[::-1]                          # (ms, mt, mu), x fastest
        self.iflag = (1 if nufft_type == 1 else -1) if isign is None else isign
        self.kp, self.nf, _, _ = orc.plan_params(nufft_type, self.modes, eps, dtype)
        self.kers = [orc.fwkerhalf(self.nf[d], self.kp) for d in range(3)]
        self.pad = slab_halo(self.kp.ns)
        self.z0, self.z1 = slab_range(self.nf[2], world, rank)
        self.nz = self.z1 - self.z0
        self.planes = np.arange(self.z0 - self.pad, self.z1 + self.pad) % self.nf[2]    # local -> global plane
        self.fw = None
        self._halo = None

    def info(self):
        return dict(z0=self.z0, z1=self.z1, pad=self.pad, nz_local=self.nz + 2 * self.pad, nf1=self.nf[0], nf2=self.nf[1],
                    nf3=self.nf[2], plane_cells=self.nf[0] * self.nf[1], rank=self.rank, world=self.world, ns=self.kp.ns)

    @staticmethod
    def _np(a):
        return a.numpy() if hasattr(a, "numpy") else np.asarray(a)

    def set_pts(self, kz, ky, kx):
        self.pts = [np.ascontiguousarray(self._np(a), self.dtype) for a in (kx, ky, kz)]

    def _fft(self, a, axes):
        f = np.fft.ifftn if self.iflag >= 0 else np.fft.fftn
        scale = float(np.prod([a.shape[i] for i in axes])) if self.iflag >= 0 else 1.0
        return (f(a, axes=axes) * scale).astype(self.cd)

    # ---- type 1 ----
    def type1_spread(self, c):
        c = self._np(c).astype(self.cd)
        if self.world == 1:
            # the halo planes alias owned planes of the same rank: attribute what wraps below plane 0
            # to the low halo and what wraps above nf3 to the high halo, as the unwrapped local grid does
            from cufinufft_b200.multi import slab_of_points
            lower = slab_of_points(self.pts[2], self.nf[2], 2) == 0
            self.fw = np.zeros((self.nz + 2 * self.pad, self.nf[1], self.nf[0]), self.cd)
            for sel, halo, wrapped in ((lower, slice(0, self.pad), slice(self.nf[2] - self.pad, self.nf[2])),
                                       (~lower, slice(self.pad + self.nz, None), slice(0, self.pad))):
                part = orc.spread([p[sel] for p in self.pts], c[sel], self.nf, self.kp)
                self.fw[halo] += part[wrapped]
                part[wrapped] = 0
                self.fw[self.pad: self.pad + self.nz] += part
            return
        full = orc.spread(self.pts, c, self.nf, self.kp)
        touched = np.zeros(self.nf[2], bool)
        touched[self.planes] = True
        assert not np.any(full[~touched]), "a point spread outside its slab + halo"
        self.fw = full[self.planes].copy()                         # [nz + 2 pad][nf2][nf1]

    def halo_pack(self, side, buf):
        src = self.fw[: self.pad] if side == 0 else self.fw[self.pad + self.nz:]
        self._np(buf)[:] = src.ravel()

    def halo_add(self, side, buf):
        b = self._np(buf).reshape(self.pad, self.nf[1], self.nf[0])
        if side == 0:
            self.fw[self.pad: 2 * self.pad] += b
        else:
            self.fw[self.nz: self.nz + self.pad] += b

    def type1_finish(self, fk):
        ms, mt, mu = self.modes
        own = self._fft(self.fw[self.pad: self.pad + self.nz], (1, 2))
        gx, gy, gz = _mode_to_grid(ms, self.nf[0]), _mode_to_grid(mt, self.nf[1]), _mode_to_grid(mu, self.nf[2])
        zbuf = np.zeros((self.nf[2], mt, ms), self.cd)
        zbuf[self.z0: self.z1] = own[:, gy][:, :, gx]
        zbuf = self._fft(zbuf, (0,))
        k1 = self.kers[0][np.abs(np.arange(ms) - ms // 2)]
        k2 = self.kers[1][np.abs(np.arange(mt) - mt // 2)]
        k3 = self.kers[2][np.abs(np.arange(mu) - mu // 2)]
        kv = (k1[None, None, :] * k2[None, :, None]) * k3[:, None, None]      # product in the real dtype, as the kernel
        out = zbuf[gz] / kv
        self._np(fk).reshape(mu, mt, ms)[:] = out.astype(self.cd)

    # ---- type 2 ----
    def type2(self, c, fk):
        ms, mt, mu = self.modes
        fk = self._np(fk).reshape(mu, mt, ms).astype(self.cd)
        gx, gy, gz = _mode_to_grid(ms, self.nf[0]), _mode_to_grid(mt, self.nf[1]), _mode_to_grid(mu, self.nf[2])
        k1 = self.kers[0][np.abs(np.arange(ms) - ms // 2)]
        k2 = self.kers[1][np.abs(np.arange(mt) - mt // 2)]
        k3 = self.kers[2][np.abs(np.arange(mu) - mu // 2)]
        kv = (k1[None, None, :] * k2[None, :, None]) * k3[:, None, None]
        zbuf = np.zeros((self.nf[2], mt, ms), self.cd)
        zbuf[gz] = (fk / kv).astype(self.cd)
        zbuf = self._fft(zbuf, (0,))
        loc = np.zeros((self.nz + 2 * self.pad, self.nf[1], self.nf[0]), self.cd)
        loc[:, gy[:, None], gx[None, :]] = zbuf[self.planes]
        loc = self._fft(loc, (1, 2))
        # interpolate on a global grid that is NaN outside this rank's planes: reading them would show
        full = np.full((self.nf[2], self.nf[1], self.nf[0]), np.nan + 1j * np.nan, self.cd)
        full[self.planes] = loc
        out = orc.interp(self.pts, full, self.nf, self.kp)
        self._np(c)[:] = out

    def halo_buffers(self):
        if self._halo is None:
            import torch
            n = self.pad * self.nf[0] * self.nf[1]
            cd = torch.complex64 if self.dtype == np.float32 else torch.complex128
            self._halo = tuple(torch.zeros(n, dtype=cd) for _ in range(4))
        return self._halo
