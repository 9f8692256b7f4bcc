"""An INDEPENDENT float64 numpy statement of the reference's GBuffer shaders (CopyTiledFramebuffer.comp, Denoise/Reproject.comp,
Denoise/Filter.comp, GBufferBlit.frag, GBuffer.h:31-130), written from the GLSL and vectorised over the image — not from
oracle/vrt_post_oracle.c.  tests/test_post_oracle.py compares the C oracle with it within a float tolerance: the GLSL cannot
run here, so two separately written restatements agreeing is the strongest statement available about the oracle's LOGIC
(indexing, weights, validity tests, buffer rotation); the oracle alone defines the rounding."""
from __future__ import annotations

import numpy as np

from scenes import gbuffer_synth as pu


def _h(x):  # rgba16f / rg16f image store
    with np.errstate(over="ignore"):
        return np.asarray(x, np.float64).astype(np.float16).astype(np.float64)


def _normal(albedo_u32):  # unpackGNormal, GBuffer.glsl:15-17
    a = (albedo_u32 >> 24).astype(np.int64)
    return np.stack([(a & 3) - 1, ((a >> 2) & 3) - 1, ((a >> 4) & 3) - 1], axis=-1).astype(np.float64)


def _luma(c):
    return c[..., 0] * 0.299 + c[..., 1] * 0.587 + c[..., 2] * 0.114


def _shift(a, dx, dy, fill=0):
    """b[y, x] = a[y + dy, x + dx] where that texel exists, else `fill`; second result = the texel exists."""
    h, w = a.shape[:2]
    out = np.full_like(a, fill)
    ok = np.zeros((h, w), bool)
    ys, ye = max(0, -dy), min(h, h - dy)
    xs, xe = max(0, -dx), min(w, w - dx)
    if ys < ye and xs < xe:
        out[ys:ye, xs:xe] = a[ys + dy : ye + dy, xs + dx : xe + dx]
        ok[ys:ye, xs:xe] = True
    return out, ok


class NumpyGBuffer:
    def __init__(self, w, h):
        self.w, self.h = w, h
        z = lambda *s, d=np.float64: np.zeros((h, w) + s, d)  # noqa: E731
        self.albedo, self.prev_albedo = z(d=np.uint32), z(d=np.uint32)
        self.irr, self.prev_irr, self.temp_irr = z(4), z(4), z(4)
        self.depth, self.prev_depth = z(), z()
        self.moments, self.prev_moments = z(2), z(2)
        self.hist = z(d=np.int64)
        self.cur = None

    def set_camera(self, proj, inv, pos):  # GBuffer::SetCamera
        m = lambda a: np.asarray(a, np.float64).reshape(4, 4).T  # noqa: E731  column-major float[16] -> M[row, col]
        cur = (m(proj), m(inv), np.asarray(pos, np.float64))
        self.history = self.cur if self.cur is not None else cur
        self.cur = cur
        self.albedo, self.prev_albedo = self.prev_albedo, self.albedo
        self.depth, self.prev_depth = self.prev_depth, self.depth
        self.moments, self.prev_moments = self.prev_moments, self.moments

    # ---- passes -----------------------------------------------------------------------------------------------------
    def _blit(self, tiles):
        w, h = self.w, self.h
        a = pu.untile(tiles, w, h, "albedo")
        d = pu.untile(tiles, w, h, "depth").astype(np.float64)
        self.albedo = np.where(d < 0, (a & 0xFF000000) | 0xFFFFFF, a).astype(np.uint32)
        rg, bx = pu.untile(tiles, w, h, "irr_rg"), pu.untile(tiles, w, h, "irr_bx")
        f = lambda u: (u & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float64)  # noqa: E731
        self.irr = np.stack([f(rg), f(rg >> 16), f(bx), np.zeros((h, w))], axis=-1)
        self.depth = d

    @staticmethod
    def _world(inv, x, y, depth):
        v = np.stack([x, y, depth, np.ones_like(depth)], axis=-1) @ inv.T
        return v[..., :3] * (16.0 / v[..., 3:4])

    def _reproject(self, reset):
        w, h = self.w, self.h
        (_, inv, pos), (hproj, hinv, hpos) = self.cur, self.history
        delta = (pos - hpos).astype(np.float32).astype(np.float64)
        X, Y = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        with np.errstate(all="ignore"):
            ok = self.depth > 0
            wp = self._world(inv, X, Y, self.depth)
            ndc = np.concatenate([wp + delta, np.ones((h, w, 1))], axis=-1) @ hproj.T
            pp = (ndc[..., :2] / ndc[..., 3:4] * 0.5 + 0.5) * np.array([w, h]) - 0.5
            pp = np.where(np.isfinite(pp), pp, -1e6)
            ppi = np.trunc(pp).astype(np.int64)
            ppf = pp - np.floor(pp)
            ok &= (ppi[..., 0] >= 0) & (ppi[..., 0] < w) & (ppi[..., 1] >= 0) & (ppi[..., 1] < h)
            cn = _normal(self.albedo)
            wsum = np.zeros((h, w))
            pirr, pmom = np.zeros((h, w, 3)), np.zeros((h, w, 2))
            hl = self.hist.copy()
            old_hist = self.hist.copy()
            pn = _normal(self.prev_albedo)
            for i in range(4):
                sx, sy = ppi[..., 0] + (i & 1), ppi[..., 1] + (i >> 1)
                inb = (sx >= 0) & (sx < w) & (sy >= 0) & (sy < h)
                cx, cy = np.clip(sx, 0, w - 1), np.clip(sy, 0, h - 1)
                sd = self.prev_depth[cy, cx]
                valid = ok & inb & ((cn * pn[cy, cx]).sum(-1) >= 0.5) & (sd > 0)
                sw = self._world(hinv, cx.astype(np.float64), cy.astype(np.float64), sd)
                valid &= ~(np.abs(((wp - sw + delta) * cn).sum(-1)) > 6.0)
                wt = np.where(i & 1, ppf[..., 0], 1 - ppf[..., 0]) * np.where(i >> 1, ppf[..., 1], 1 - ppf[..., 1])
                wt = np.where(valid, wt, 0.0)
                pirr += self.prev_irr[cy, cx, :3] * wt[..., None]
                pmom += self.prev_moments[cy, cx] * wt[..., None]
                wsum += wt
                hl = np.where(valid, np.minimum(hl, old_hist[cy, cx] + 1), hl)
            ok &= ~(wsum < 0.001)
            pirr, pmom = pirr / wsum[..., None], pmom / wsum[..., None]
            if reset:
                hl = np.minimum(hl, 6)
            blend = 1.0 / (hl + 1)
            nirr = pirr * (1 - blend[..., None]) + self.irr[..., :3] * blend[..., None]
            luma = _luma(nirr)
            mb = np.maximum(0.5, blend)[..., None]
            nmom = pmom * (1 - mb) + np.stack([luma, luma * luma], -1) * mb
            var = np.maximum(0.0, nmom[..., 1] - nmom[..., 0] ** 2)
        o3, o2 = ok[..., None], ok[..., None]
        self.moments = np.where(o2, _h(nmom), 0.0)
        self.irr = np.where(o3, _h(np.concatenate([nirr, var[..., None]], -1)), self.irr)
        self.hist = np.where(ok, np.minimum(hl + 1, 64), 0)

    def _tap_weight(self, w_luma, cd, d_tap, length, n_tap, cn):
        w_normal = np.clip((n_tap * cn).sum(-1), 0.001, 1.0) ** 128
        w_depth = np.abs(cd - d_tap) / (length + 0.001)
        return np.exp(-(w_luma + w_depth)) * w_normal

    def _variance(self):
        skip = (self.hist > 4) | (self.depth < 0)
        cl = _luma(self.temp_irr)  # Filter.comp:26: the centre comes from the TEMP texture
        cn = _normal(self.albedo)
        nrm = cn
        si, sm, wsum = np.zeros((self.h, self.w, 3)), np.zeros((self.h, self.w, 2)), np.zeros((self.h, self.w))
        for ky in range(-3, 4):
            for kx in range(-3, 4):
                irr, ok = _shift(self.irr, kx, ky)
                n_tap, _ = _shift(nrm, kx, ky)
                d_tap, _ = _shift(self.depth, kx, ky)
                l = _luma(irr)
                wt = self._tap_weight(np.abs(l - cl) / 10.0, self.depth, d_tap, np.hypot(kx, ky), n_tap, cn)
                wt = np.where(ok, wt, 0.0)
                si += irr[..., :3] * wt[..., None]
                sm += np.stack([l, l * l], -1) * wt[..., None]
                wsum += wt
        wsum = np.maximum(wsum, 0.001)
        si, sm = si / wsum[..., None], sm / wsum[..., None]
        var = np.maximum(0.0, sm[..., 1] - sm[..., 0] ** 2) * (4.0 - self.hist) * 3.0
        out = _h(np.concatenate([si, var[..., None]], -1))
        self.temp_irr = np.where(skip[..., None], self.irr, out)

    def _atrous(self, src, pass_no):
        kvar = np.array([[1 / 4, 1 / 8], [1 / 8, 1 / 16]])
        kern = [3 / 8, 1 / 4, 1 / 16]
        cv = np.zeros((self.h, self.w))
        for ky in (-1, 0, 1):
            for kx in (-1, 0, 1):
                v, _ = _shift(src[..., 3], kx, ky)  # imageLoad outside the image: 0
                cv += v * kvar[abs(kx)][abs(ky)]
        cn = _normal(self.albedo)
        cl = _luma(src)
        phi = np.sqrt(np.maximum(0.0001, cv)) * 4.0
        total, wsum = src.copy(), np.ones((self.h, self.w))
        for ky in range(-2, 3):
            for kx in range(-2, 3):
                if kx == 0 and ky == 0:
                    continue
                ox, oy = kx << pass_no, ky << pass_no
                irr, ok = _shift(src, ox, oy)
                n_tap, _ = _shift(cn, ox, oy)
                d_tap, _ = _shift(self.depth, ox, oy)
                wt = kern[abs(kx)] * kern[abs(ky)] * self._tap_weight(np.abs(_luma(irr) - cl) / phi, self.depth, d_tap, np.hypot(ox, oy), n_tap, cn)
                wt = np.where(ok, wt, 0.0)
                total[..., :3] += irr[..., :3] * wt[..., None]
                total[..., 3] += irr[..., 3] * wt * wt
                wsum += wt
        wsum = np.maximum(wsum, 0.001)
        total[..., :3] /= wsum[..., None]
        total[..., 3] /= wsum * wsum
        return np.where((self.depth < 0)[..., None], src, _h(total))

    def _present(self):
        alb = np.stack([(self.albedo >> s) & 255 for s in (0, 8, 16)], -1) / 255.0
        v = alb * self.irr[..., :3] * 0.48 * 0.6
        with np.errstate(all="ignore"):
            c = np.clip((v * (2.51 * v + 0.03)) / (v * (2.43 * v + 0.59) + 0.14), 0, 1) ** 0.45
        return np.floor(np.nan_to_num(c) * 255 + 0.5).astype(np.int64)

    def frame(self, tiles, reset=False, passes=5):  # CopyTiledFramebuffer + GBuffer::DenoiseAndPresent (GBuffer.h:86-130)
        self._blit(tiles)
        self._reproject(reset)
        if passes > 0:
            self._variance()
            for i in range(passes):
                src = self.prev_irr if i == 1 else (self.temp_irr if i % 2 == 0 else self.irr)
                out = self._atrous(src, i)
                if i % 2 == 0:
                    self.irr = out
                else:
                    self.temp_irr = out
                if i == 0:
                    self.prev_irr, self.irr = self.irr, self.prev_irr
            if passes % 2 != 0:
                self.temp_irr, self.irr = self.irr, self.temp_irr
        rgb = self._present()
        if passes == 0:
            self.prev_irr, self.irr = self.irr, self.prev_irr
        return rgb
