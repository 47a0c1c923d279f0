"""TEST INFRASTRUCTURE — CPU restatement of the JPEG decode that the reference gets from `cv2.imread`
(vltk/compat.py:573-579 `img_tensorize`, called at legacy/processing.py:119-129).  Never imported by the product.

The decoder itself is a third-party dependency that is NOT in the reference tree: libjpeg-turbo, bundled inside
the opencv-python / Pillow wheels (requirements.txt names `opencv-python` and `Pillow` without pins; this container
has opencv 4.13.0 with libjpeg-turbo 3.1.2 and Pillow 12.2.0).  What is restated here is its published default
pipeline — ITU-T T.81 Huffman decoding (Annex F.2), the IJG "islow" integer inverse DCT (jidctint.c: 13-bit
constants, PASS1_BITS 2), "fancy" triangle-filter chroma upsampling (jdsample.c h2v1/h2v2_fancy_upsample, edge
rows/columns replicated) and the 16-bit fixed-point YCbCr->RGB tables (jdcolor.c).  All of it is integer
arithmetic, so parity is BIT-EXACT.

Pinned: tests/test_jpeg.py checks `decode()` against `cv2.imdecode` (and PIL) on every fixture, here and on the
GPU box (both ship the same wheels).  Pure-Python entropy decoding: use small images.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
                   7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                   39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


# ------------------------------------------------------------------ markers (T.81 Annex B)
def parse(data: bytes):
    assert data[:2] == b"\xff\xd8", "no SOI"
    pos, qt, dht, info = 2, {}, {}, {"dri": 0, "adobe": -1}
    while pos < len(data):
        assert data[pos] == 0xFF
        while data[pos] == 0xFF:
            pos += 1
        m = data[pos]
        pos += 1
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        if m == 0xD9:
            break
        ln = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + ln]
        if m == 0xDB:
            o = 0
            while o < len(seg):
                pq, t = seg[o] >> 4, seg[o] & 15
                o += 1
                if pq:
                    vals = [(seg[o + 2 * i] << 8) | seg[o + 2 * i + 1] for i in range(64)]
                    o += 128
                else:
                    vals = list(seg[o:o + 64])
                    o += 64
                nat = np.zeros(64, np.int64)
                nat[ZIGZAG] = vals
                qt[t] = nat
        elif m == 0xC4:
            o = 0
            while o < len(seg):
                tc, th = seg[o] >> 4, seg[o] & 15
                bits = list(seg[o + 1:o + 17])
                o += 17
                n = sum(bits)
                dht[(tc, th)] = _huff_table(bits, list(seg[o:o + n]))
                o += n
        elif m in (0xC0, 0xC1, 0xC2):
            info["progressive"] = m == 0xC2
            assert seg[0] == 8
            info["h"], info["w"] = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            info["comps"] = [dict(id=seg[6 + 3 * c], hs=seg[7 + 3 * c] >> 4, vs=seg[7 + 3 * c] & 15, tq=seg[8 + 3 * c])
                             for c in range(seg[5])]
        elif 0xC3 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise NotImplementedError(f"SOF{m - 0xC0}")
        elif m == 0xDD:
            info["dri"] = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            info["adobe"] = seg[11]
        elif m == 0xDA:
            if info.get("progressive"):     # geometry + tables only: the scans are not restated in Python
                info["scan"] = pos - 2
                break
            ns = seg[0]
            assert ns == len(info["comps"])
            for k in range(ns):
                assert seg[1 + 2 * k] == info["comps"][k]["id"]
                info["comps"][k]["td"], info["comps"][k]["ta"] = seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15
            info["scan"] = pos + ln
            break
        pos += ln
    info["qt"], info["dht"] = qt, dht
    return info


def _huff_table(bits, vals):
    """code -> symbol dictionary keyed by (length, code) — the canonical assignment of T.81 Annex C."""
    table, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            table[(ln, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


# ------------------------------------------------------------------ entropy decoding (T.81 F.2)
class _Bits:
    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def bit(self):
        if self.n == 0:
            b = self.d[self.p] if self.p < len(self.d) else 0
            if b == 0xFF:
                nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                if nxt == 0:
                    self.p += 2
                else:
                    b = 0           # marker: feed zeros, do not advance
            else:
                self.p += 1
            self.acc, self.n = b, 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, s):
        v = 0
        for _ in range(s):
            v = (v << 1) | self.bit()
        return v

    def huff(self, table):
        code = 0
        for ln in range(1, 17):
            code = (code << 1) | self.bit()
            if (ln, code) in table:
                return table[(ln, code)]
        raise ValueError("bad Huffman code")

    def restart(self):
        self.n = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _extend(v, s):
    return v - (1 << s) + 1 if s and v < (1 << (s - 1)) else v


def coefficients(data: bytes):
    """-> (info, [per-component int16 array [blocks_h, blocks_w, 64] in natural order])."""
    I = parse(data)
    comps = I["comps"]
    if len(comps) == 1:
        comps[0]["hs"] = comps[0]["vs"] = 1
    hmax, vmax = max(c["hs"] for c in comps), max(c["vs"] for c in comps)
    mx, my = -(-I["w"] // (8 * hmax)), -(-I["h"] // (8 * vmax))
    out = [np.zeros((my * c["vs"], mx * c["hs"], 64), np.int16) for c in comps]
    br, pred, left = _Bits(data, I["scan"]), [0] * len(comps), I["dri"]
    for y in range(my):
        for x in range(mx):
            if I["dri"] and left == 0:
                br.restart()
                pred, left = [0] * len(comps), I["dri"]
            for ci, c in enumerate(comps):
                dc, ac = I["dht"][(0, c["td"])], I["dht"][(1, c["ta"])]
                for by in range(c["vs"]):
                    for bx in range(c["hs"]):
                        blk = out[ci][y * c["vs"] + by, x * c["hs"] + bx]
                        s = br.huff(dc)
                        pred[ci] += _extend(br.bits(s), s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = br.huff(ac)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                            k += 1
            left -= 1
    I.update(hmax=hmax, vmax=vmax, mcus_x=mx, mcus_y=my)
    return I, out


# ------------------------------------------------------------------ islow inverse DCT (jidctint.c)
_C = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137,
          f1961=16069, f2053=16819, f2562=20995, f3072=25172)


def _idct_1d(v, shift):
    """v: [..., 8] int64 along the last axis -> descaled outputs [..., 8]."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * _C["f0541"]
    tmp2 = z1 - z3 * _C["f1847"]
    tmp3 = z1 + z2 * _C["f0765"]
    tmp0 = (v[..., 0] + v[..., 4]) << 13
    tmp1 = (v[..., 0] - v[..., 4]) << 13
    t10, t13, t11, t12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * _C["f1175"]
    tmp0, tmp1, tmp2, tmp3 = tmp0 * _C["f0298"], tmp1 * _C["f2053"], tmp2 * _C["f3072"], tmp3 * _C["f1501"]
    z1, z2, z3, z4 = -z1 * _C["f0899"], -z2 * _C["f2562"], -z3 * _C["f1961"] + z5, -z4 * _C["f0390"] + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    res = np.stack([t10 + tmp3, t11 + tmp2, t12 + tmp1, t13 + tmp0, t13 - tmp0, t12 - tmp1, t11 - tmp2, t10 - tmp3], -1)
    return (res + (1 << (shift - 1))) >> shift


def idct_plane(coef, qt):
    """coef [bh, bw, 64] quantised, qt [64] -> u8 plane [bh*8, bw*8]."""
    bh, bw, _ = coef.shape
    x = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(bh, bw, 8, 8)
    ws = _idct_1d(np.swapaxes(x, -1, -2), 13 - 2)           # pass 1: columns
    ws = np.swapaxes(ws, -1, -2)
    px = _idct_1d(ws, 13 + 2 + 3)                           # pass 2: rows
    px = np.clip(px + 128, 0, 255).astype(np.uint8)
    return px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


# ------------------------------------------------------------------ upsampling (jdsample.c)
def _h2_fancy(row_near3_plus_far, bias_even, bias_odd, shift):
    """Horizontal triangle filter on column sums cs [..., W] -> [..., 2W]."""
    cs = row_near3_plus_far.astype(np.int64)
    prev = np.concatenate([cs[..., :1], cs[..., :-1]], -1)
    nxt = np.concatenate([cs[..., 1:], cs[..., -1:]], -1)
    out = np.empty(cs.shape[:-1] + (cs.shape[-1] * 2,), np.int64)
    out[..., 0::2] = (cs * 3 + prev + bias_even) >> shift
    out[..., 1::2] = (cs * 3 + nxt + bias_odd) >> shift
    return out


def upsample(plane, cw, ch, h2, v2, W, H):
    """Chroma plane (real extent cw x ch inside a padded plane) -> [H, W] at luma resolution."""
    p = plane[:ch, :cw].astype(np.int64)
    if not h2:
        return p[:H, :W]
    if cw <= 2:                                              # too narrow for the fancy filter: replication
        p = np.repeat(p, 2, 1)
        return (np.repeat(p, 2, 0) if v2 else p)[:H, :W]
    if v2:
        above = np.concatenate([p[:1], p[:-1]], 0)
        below = np.concatenate([p[1:], p[-1:]], 0)
        rows = np.empty((2 * ch, cw), np.int64)
        rows[0::2] = p * 3 + above                           # v == 0: next nearest is the row above
        rows[1::2] = p * 3 + below
        return _h2_fancy(rows, 8, 7, 4)[:H, :W]
    return _h2_fancy(p, 1, 2, 2)[:H, :W]                     # h2v1: (3*in + neighbour + 1|2) >> 2


# ------------------------------------------------------------------ colour (jdcolor.c)
def ycc_to_bgr(Y, Cb, Cr):
    Y, cb, cr = Y.astype(np.int64), Cb.astype(np.int64) - 128, Cr.astype(np.int64) - 128
    r = Y + ((91881 * cr + 32768) >> 16)
    b = Y + ((116130 * cb + 32768) >> 16)
    g = Y + ((-22554 * cb - 46802 * cr + 32768) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def geometry(I):
    """Adds hmax/vmax/mcus_x/mcus_y to a parsed header (what coefficients() does for baseline files)."""
    comps = I["comps"]
    if len(comps) == 1:
        comps[0]["hs"] = comps[0]["vs"] = 1
    hmax, vmax = max(c["hs"] for c in comps), max(c["vs"] for c in comps)
    I.update(hmax=hmax, vmax=vmax, mcus_x=-(-I["w"] // (8 * hmax)), mcus_y=-(-I["h"] // (8 * vmax)))
    return I


def reconstruct(I, coefs):
    comps, W, H = I["comps"], I["w"], I["h"]
    planes = [idct_plane(coefs[i], I["qt"][c["tq"]]) for i, c in enumerate(comps)]
    if len(comps) == 1:
        g = planes[0][:H, :W]
        return np.stack([g, g, g], -1)
    h2, v2 = comps[0]["hs"] == 2, comps[0]["vs"] == 2
    cw, ch = -(-W * comps[1]["hs"] // I["hmax"]), -(-H * comps[1]["vs"] // I["vmax"])
    Y = planes[0][:H, :W]
    c1 = upsample(planes[1], cw, ch, h2, v2, W, H)
    c2 = upsample(planes[2], cw, ch, h2, v2, W, H)
    if I["adobe"] == 0 or [c["id"] for c in comps] == [82, 71, 66]:
        return np.stack([c2, c1, Y], -1).astype(np.uint8)
    return ycc_to_bgr(Y, c1, c2)


def decode(data: bytes) -> np.ndarray:
    """JPEG bytes -> BGR u8 [h, w, 3], what cv2.imdecode(..., IMREAD_COLOR | IMREAD_IGNORE_ORIENTATION) returns."""
    I, coefs = coefficients(data)
    return reconstruct(I, coefs)
