"""Oracle: baseline JPEG decode, bit for bit as libjpeg-turbo does it (TEST INFRASTRUCTURE).

The reference loads images with ``spacer.storage.load_image`` -> ``PIL.Image.open(...).convert("RGB")`` (call site
``mermaid_classifier/pyspacer/annotation.py:235``; inside ``spacer.tasks.extract_features``,
``scripts/build_feature_bucket.py:775``); PIL decodes JPEG through libjpeg-turbo with its defaults: the accurate integer
inverse DCT (``jidctint.c`` ``jpeg_idct_islow``), "fancy" (triangle) chroma upsampling (``jdsample.c``) and the fixed-point
YCbCr -> RGB tables of ``jdcolor.c``.  This module restates that pipeline -- entropy decoding in plain Python (small images
only), everything after it in NumPy integer arithmetic -- and ``tests/test_oracle_jpeg.py`` pins it to PIL itself, byte for
byte, on generated fixtures (4:4:4, 4:2:2, 4:2:0, grayscale, odd sizes, restart intervals, low quality).

Covered: baseline / extended-sequential and progressive Huffman streams, 8-bit, 1 or 3 components, chroma sampling 1x1 with
luma 1x1 / 2x1 / 2x2.  Anything else (arithmetic coding, CMYK, 4:4:0, 4:1:1 ...) raises ``UnsupportedJpeg``: the product
falls back too.
"""

from __future__ import annotations

import numpy as np


class UnsupportedJpeg(ValueError):
    pass


ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46,
                   53, 60, 61, 54, 47, 55, 62, 63], dtype=np.int64)


# ---------------------------------------------------------------------------------------------------------------------
# header parsing
# ---------------------------------------------------------------------------------------------------------------------
def parse(data: bytes) -> dict:
    """Markers of a baseline or progressive JPEG: quantisation tables (natural order), frame header, restart interval, and
    every scan with the Huffman tables in force when it starts and the offset of its entropy-coded segment."""
    if data[:2] != b"\xff\xd8":
        raise UnsupportedJpeg("not a JPEG stream")
    qt: dict[int, np.ndarray] = {}
    ht: dict[tuple[int, int], tuple[list[int], list[int]]] = {}
    frame = None
    ri = 0
    adobe_transform = None
    scans = []
    pos = 2
    n = len(data)
    while pos < n:
        if data[pos] != 0xFF:
            raise UnsupportedJpeg("marker expected")
        while pos < n and data[pos] == 0xFF:
            pos += 1
        m = data[pos]
        pos += 1
        if m in (0x01,) or 0xD0 <= m <= 0xD7:
            continue
        if m == 0xD9:
            break
        seg_len = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2: pos + seg_len]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    vals = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    vals = list(seg[i: i + 64])
                    i += 64
                t = np.zeros(64, dtype=np.int64)
                t[ZIGZAG] = vals
                qt[tq] = t
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1: i + 17])
                k = sum(counts)
                ht[(tc, th)] = (counts, list(seg[i + 17: i + 17 + k]))
                i += 17 + k
        elif m in (0xC0, 0xC1, 0xC2):
            if seg[0] != 8:
                raise UnsupportedJpeg("sample precision")
            h, w, nf = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [{"id": seg[6 + 3 * c], "h": seg[7 + 3 * c] >> 4, "v": seg[7 + 3 * c] & 15, "tq": seg[8 + 3 * c]} for c in range(nf)]
            frame = {"height": h, "width": w, "comps": comps, "progressive": m == 0xC2}
        elif 0xC3 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise UnsupportedJpeg("not a Huffman-coded baseline / progressive JPEG")
        elif m == 0xDD:
            ri = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe_transform = seg[11]
        elif m == 0xDA:
            if frame is None:
                raise UnsupportedJpeg("scan before frame")
            ns = seg[0]
            ids = [c["id"] for c in frame["comps"]]
            sc = [{"ci": ids.index(seg[1 + 2 * c]), "td": seg[2 + 2 * c] >> 4, "ta": seg[2 + 2 * c] & 15} for c in range(ns)]
            ss, se, ahal = seg[1 + 2 * ns], seg[2 + 2 * ns], seg[3 + 2 * ns]
            scans.append({"comps": sc, "ss": ss, "se": se, "ah": ahal >> 4, "al": ahal & 15, "ht": dict(ht), "ri": ri,
                          "data_pos": pos + seg_len})
            # skip the entropy-coded segment: up to the next marker that is not RSTn / stuffing
            q = pos + seg_len
            while q + 1 < n and not (data[q] == 0xFF and data[q + 1] != 0 and not (0xD0 <= data[q + 1] <= 0xD7)):
                q += 1
            pos = q
            continue
        pos += seg_len
    if frame is None or not scans:
        raise UnsupportedJpeg("no scan")
    if not frame["progressive"] and (len(scans) != 1 or len(scans[0]["comps"]) != len(frame["comps"])):
        raise UnsupportedJpeg("multi-scan sequential files")
    first = scans[0]
    return {"qt": qt, "ht": first["ht"], "frame": frame, "scan": [{"id": frame["comps"][c["ci"]]["id"], "td": c["td"], "ta": c["ta"]} for c in first["comps"]],
            "scans": scans, "ri": first["ri"], "data_pos": first["data_pos"], "adobe_transform": adobe_transform}


def check_supported(hdr: dict) -> None:
    comps = hdr["frame"]["comps"]
    if len(comps) == 1:
        return
    if len(comps) != 3 or hdr["adobe_transform"] == 0:
        raise UnsupportedJpeg("colour space")
    if (comps[1]["h"], comps[1]["v"], comps[2]["h"], comps[2]["v"]) != (1, 1, 1, 1):
        raise UnsupportedJpeg("chroma sampling factors")
    if (comps[0]["h"], comps[0]["v"]) not in ((1, 1), (2, 1), (2, 2)):
        raise UnsupportedJpeg("luma sampling factors")


# ---------------------------------------------------------------------------------------------------------------------
# entropy decoding (plain Python: small images only)
# ---------------------------------------------------------------------------------------------------------------------
class _Huff:
    def __init__(self, counts, symbols):
        self.look = {}
        code = 0
        k = 0
        for length in range(1, 17):
            for _ in range(counts[length - 1]):
                self.look[(length, code)] = symbols[k]
                code += 1
                k += 1
            code <<= 1


class _Bits:
    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def bit(self) -> int:
        if self.n == 0:
            b = self.d[self.p]
            self.p += 1
            if b == 0xFF:
                nxt = self.d[self.p]
                if nxt == 0:
                    self.p += 1
                else:   # a marker inside the data: feed zeros (libjpeg does the same with a warning)
                    self.p -= 1
                    b = 0
            self.acc, self.n = b, 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k: int) -> int:
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self) -> None:
        self.n = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _sym(br: _Bits, h: _Huff) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        s = h.look.get((length, code))
        if s is not None:
            return s
    raise UnsupportedJpeg("bad Huffman code")


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def decode_coefficients(data: bytes, hdr: dict) -> list[np.ndarray]:
    """Quantised DCT coefficients per component: ``(blocks_y, blocks_x, 64) int32`` in natural (row-major) order, over the
    whole MCU grid (blocks past the image edge included)."""
    fr = hdr["frame"]
    comps = fr["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    if len(comps) == 1:   # a single-component scan is not interleaved: one block per MCU, the component's own block grid
        mcux = (fr["width"] + 7) // 8
        mcuy = (fr["height"] + 7) // 8
        shape = [(mcuy, mcux)]
        per = [(1, 1)]
    else:
        mcux = (fr["width"] + 8 * hmax - 1) // (8 * hmax)
        mcuy = (fr["height"] + 8 * vmax - 1) // (8 * vmax)
        shape = [(mcuy * c["v"], mcux * c["h"]) for c in comps]
        per = [(c["v"], c["h"]) for c in comps]
    out = [np.zeros((s[0], s[1], 64), dtype=np.int32) for s in shape]
    tabs = [(_Huff(*hdr["ht"][(0, sc["td"])]), _Huff(*hdr["ht"][(1, sc["ta"])])) for sc in hdr["scan"]]
    br = _Bits(data, hdr["data_pos"])
    pred = [0] * len(comps)
    ri = hdr["ri"]
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if ri and count and count % ri == 0:
                br.restart()
                pred = [0] * len(comps)
            count += 1
            for ci in range(len(comps)):
                dc_t, ac_t = tabs[ci]
                v, h = per[ci]
                for by in range(v):
                    for bx in range(h):
                        blk = out[ci][my * v + by, mx * h + bx]
                        s = _sym(br, dc_t)
                        diff = _extend(br.bits(s), s) if s else 0
                        pred[ci] += diff
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _sym(br, ac_t)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                            k += 1
    return out


def decode_coefficients_progressive(data: bytes, hdr: dict) -> list[np.ndarray]:
    """``jdphuff.c``: DC / AC first and refinement scans accumulated into the coefficient arrays (same layout as
    :func:`decode_coefficients`)."""
    fr = hdr["frame"]
    comps = fr["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    W, H = fr["width"], fr["height"]
    if len(comps) == 1:
        hmax = vmax = 1
        grid = [((H + 7) // 8, (W + 7) // 8)]
    else:
        mcux = (W + 8 * hmax - 1) // (8 * hmax)
        mcuy = (H + 8 * vmax - 1) // (8 * vmax)
        grid = [(mcuy * c["v"], mcux * c["h"]) for c in comps]
    out = [np.zeros((g[0], g[1], 64), dtype=np.int32) for g in grid]
    for sc in hdr["scans"]:
        br = _Bits(data, sc["data_pos"])
        ss, se, ah, al = sc["ss"], sc["se"], sc["ah"], sc["al"]
        ns = len(sc["comps"])
        tabs = {}
        for c in sc["comps"]:
            tabs[c["ci"]] = (_Huff(*sc["ht"][(0, c["td"])]) if (0, c["td"]) in sc["ht"] else None,
                             _Huff(*sc["ht"][(1, c["ta"])]) if (1, c["ta"]) in sc["ht"] else None)
        # block visiting order of this scan
        if ns > 1:
            order = []
            for my in range(mcuy):
                for mx in range(mcux):
                    unit = []
                    for c in sc["comps"]:
                        ci = c["ci"]
                        v, h = (comps[ci]["v"], comps[ci]["h"]) if len(comps) > 1 else (1, 1)
                        unit += [(ci, my * v + by, mx * h + bx) for by in range(v) for bx in range(h)]
                    order.append(unit)
        else:
            ci = sc["comps"][0]["ci"]
            ch_, cw_ = comps[ci]["h"], comps[ci]["v"]
            if len(comps) == 1:
                bw, bh = (W + 7) // 8, (H + 7) // 8
            else:
                bw = ((W * comps[ci]["h"] + hmax - 1) // hmax + 7) // 8
                bh = ((H * comps[ci]["v"] + vmax - 1) // vmax + 7) // 8
            order = [[(ci, y, x)] for y in range(bh) for x in range(bw)]
        pred = {c["ci"]: 0 for c in sc["comps"]}
        eobrun = 0
        ri = sc["ri"]
        p1, m1 = 1 << al, -1 << al
        for count, unit in enumerate(order):
            if ri and count and count % ri == 0:
                br.restart()
                pred = {k: 0 for k in pred}
                eobrun = 0
            for ci, y, x in unit:
                blk = out[ci][y, x]
                dc_t, ac_t = tabs[ci]
                if ss == 0:
                    if ah == 0:
                        s = _sym(br, dc_t)
                        pred[ci] += _extend(br.bits(s), s) if s else 0
                        blk[0] = pred[ci] << al
                    elif br.bit():
                        blk[0] |= p1
                    continue
                if ah == 0:   # AC first
                    if eobrun > 0:
                        eobrun -= 1
                        continue
                    k = ss
                    while k <= se:
                        rs = _sym(br, ac_t)
                        r, s = rs >> 4, rs & 15
                        if s:
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s) << al
                        elif r == 15:
                            k += 15
                        else:
                            eobrun = (1 << r) + (br.bits(r) if r else 0) - 1
                            break
                        k += 1
                    continue
                # AC refinement
                k = ss
                if eobrun == 0:
                    while k <= se:
                        rs = _sym(br, ac_t)
                        r, s = rs >> 4, rs & 15
                        if s:
                            s = p1 if br.bit() else m1
                        elif r != 15:
                            eobrun = (1 << r) + (br.bits(r) if r else 0)
                            break
                        while k <= se:
                            z = ZIGZAG[k]
                            if blk[z] != 0:
                                if br.bit() and (blk[z] & p1) == 0:
                                    blk[z] += p1 if blk[z] >= 0 else m1
                            else:
                                r -= 1
                                if r < 0:
                                    break
                            k += 1
                        if s:
                            blk[ZIGZAG[k]] = s
                        k += 1
                if eobrun > 0:
                    while k <= se:
                        z = ZIGZAG[k]
                        if blk[z] != 0 and br.bit() and (blk[z] & p1) == 0:
                            blk[z] += p1 if blk[z] >= 0 else m1
                        k += 1
                    eobrun -= 1
    return out


# ---------------------------------------------------------------------------------------------------------------------
# jidctint.c: jpeg_idct_islow
# ---------------------------------------------------------------------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(i0, i1, i2, i3, i4, i5, i6, i7, shift):
    """One pass of the LL&M butterfly on eight int64 arrays (no zero-AC shortcut: it is arithmetically identical)."""
    z2, z3 = i2, i6
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    tmp0 = (i0 + i4) << CONST_BITS
    tmp1 = (i0 - i4) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = i7, i5, i3, i1
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175875602
    tmp0 = tmp0 * F_0_298631336
    tmp1 = tmp1 * F_2_053119869
    tmp2 = tmp2 * F_3_072711026
    tmp3 = tmp3 * F_1_501321110
    z1 = z1 * (-F_0_899976223)
    z2 = z2 * (-F_2_562915447)
    z3 = z3 * (-F_1_961570560) + z5
    z4 = z4 * (-F_0_390180644) + z5
    tmp0 = tmp0 + z1 + z3
    tmp1 = tmp1 + z2 + z4
    tmp2 = tmp2 + z2 + z3
    tmp3 = tmp3 + z1 + z4
    return [_descale(tmp10 + tmp3, shift), _descale(tmp11 + tmp2, shift), _descale(tmp12 + tmp1, shift),
            _descale(tmp13 + tmp0, shift), _descale(tmp13 - tmp0, shift), _descale(tmp12 - tmp1, shift),
            _descale(tmp11 - tmp2, shift), _descale(tmp10 - tmp3, shift)]


def range_limit_idct(x: np.ndarray) -> np.ndarray:
    """``range_limit[x & RANGE_MASK]`` of the IDCT output (table of ``jdmaster.c:prepare_range_limit_table``): ``x + 128``
    clamped to 0..255 for ``x`` in [-512, 511], with the table's wrap-around outside (corrupt data only)."""
    i = np.asarray(x) & 1023
    out = np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896)))
    return out.astype(np.uint8)


def idct_islow(coef: np.ndarray, quant: np.ndarray) -> np.ndarray:
    """``(..., 64)`` quantised coefficients x ``(64,)`` table -> ``(..., 8, 8) uint8`` samples."""
    c = coef.astype(np.int64) * quant.astype(np.int64)
    c = c.reshape(c.shape[:-1] + (8, 8))
    rows = [c[..., r, :] for r in range(8)]                       # pass 1 works down the columns: inputs = the eight rows
    ws = _idct_1d(*rows, CONST_BITS - PASS1_BITS)                  # ws[r][..., col]
    ws = np.stack(ws, axis=-2)                                     # (..., 8 rows, 8 cols)
    cols = [ws[..., :, k] for k in range(8)]                      # pass 2 works along the rows: inputs = the eight columns
    outc = _idct_1d(*cols, CONST_BITS + PASS1_BITS + 3)
    out = np.stack(outc, axis=-1)                                  # (..., 8 rows, 8 cols)
    return range_limit_idct(out)


# ---------------------------------------------------------------------------------------------------------------------
# jdsample.c: fancy upsampling
# ---------------------------------------------------------------------------------------------------------------------
def h2v1_fancy(x: np.ndarray) -> np.ndarray:
    """``(rows, w)`` -> ``(rows, 2 w)``."""
    v = x.astype(np.int32)
    rows, w = v.shape
    out = np.empty((rows, 2 * w), dtype=np.int32)
    if w == 1:
        out[:, 0] = out[:, 1] = v[:, 0]
        return out.astype(np.uint8)
    left = np.concatenate([v[:, :1], v[:, :-1]], axis=1)
    right = np.concatenate([v[:, 1:], v[:, -1:]], axis=1)
    out[:, 0::2] = (3 * v + left + 1) >> 2
    out[:, 1::2] = (3 * v + right + 2) >> 2
    out[:, 0] = v[:, 0]
    out[:, -1] = v[:, -1]
    return out.astype(np.uint8)


def h2v2_fancy(x: np.ndarray) -> np.ndarray:
    """``(h, w)`` -> ``(2 h, 2 w)``; the rows above the first and below the last are the edge rows themselves."""
    v = x.astype(np.int32)
    h, w = v.shape
    above = np.concatenate([v[:1], v[:-1]], axis=0)
    below = np.concatenate([v[1:], v[-1:]], axis=0)
    out = np.empty((2 * h, 2 * w), dtype=np.int32)
    for k, other in ((0, above), (1, below)):
        colsum = 3 * v + other                                    # thiscolsum of every column
        if w == 1:
            out[k::2, 0] = (colsum[:, 0] * 4 + 8) >> 4
            out[k::2, 1] = (colsum[:, 0] * 4 + 7) >> 4
            continue
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
        even = (colsum * 3 + last + 8) >> 4
        odd = (colsum * 3 + nxt + 7) >> 4
        even[:, 0] = (colsum[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (colsum[:, -1] * 4 + 7) >> 4
        out[k::2, 0::2] = even
        out[k::2, 1::2] = odd
    return out.astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# jdcolor.c: YCbCr -> RGB
# ---------------------------------------------------------------------------------------------------------------------
def _fix(x: float) -> int:
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + 32768) >> 16
CB_B = (_fix(1.77200) * _X + 32768) >> 16
CR_G = -_fix(0.71414) * _X
CB_G = -_fix(0.34414) * _X + 32768


def ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    yy = y.astype(np.int64)
    r = yy + CR_R[cr]
    g = yy + ((CB_G[cb] + CR_G[cr]) >> 16)
    b = yy + CB_B[cb]
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# the whole decode
# ---------------------------------------------------------------------------------------------------------------------
def planes_from_coefficients(coefs: list[np.ndarray], hdr: dict) -> list[np.ndarray]:
    """Component planes (IDCT output of every block, assembled), each over its whole block grid."""
    planes = []
    for ci, c in enumerate(hdr["frame"]["comps"]):
        s = idct_islow(coefs[ci], hdr["qt"][c["tq"]])              # (by, bx, 8, 8)
        by, bx = s.shape[:2]
        planes.append(s.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8))
    return planes


def rgb_from_planes(planes: list[np.ndarray], hdr: dict) -> np.ndarray:
    fr = hdr["frame"]
    H, W = fr["height"], fr["width"]
    comps = fr["comps"]
    if len(comps) == 1:
        y = planes[0][:H, :W]
        return np.stack([y, y, y], axis=-1)
    hmax, vmax = comps[0]["h"], comps[0]["v"]
    y = planes[0][:H, :W]
    ch = (H * 1 + vmax - 1) // vmax            # downsampled_height / width of the chroma components
    cw = (W * 1 + hmax - 1) // hmax
    up = []
    for p in planes[1:]:
        c = p[:ch, :cw]
        if (hmax, vmax) == (1, 1):
            u = c
        elif cw <= 2:   # jdsample.c: the fancy routines need downsampled_width > 2; narrower components are replicated
            u = np.repeat(np.repeat(c, vmax, axis=0), hmax, axis=1)
        elif (hmax, vmax) == (2, 1):
            u = h2v1_fancy(c)
        else:
            u = h2v2_fancy(c)
        up.append(u[:H, :W])
    return ycc_to_rgb(y, up[0], up[1])


def decode_rgb(data: bytes) -> np.ndarray:
    """``(H, W, 3) uint8``: what ``PIL.Image.open(io.BytesIO(data)).convert("RGB")`` returns."""
    hdr = parse(data)
    check_supported(hdr)
    coefs = decode_coefficients_progressive(data, hdr) if hdr["frame"]["progressive"] else decode_coefficients(data, hdr)
    return rgb_from_planes(planes_from_coefficients(coefs, hdr), hdr)
