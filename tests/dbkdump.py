"""Reader of the deblocking dumps written by the instrumented reference decoder (oracle/dbk_dump.inc; test infrastructure).
-> list of pictures, each a dict: w, h, bit depths, slice / PPS offsets, planes before ('pre') and after ('post') deblocking,
and per 4x4 luma unit of the picture: bs_ver, bs_hor (boundary strength of the unit's left / top edge), qp, nofilter."""
import numpy as np


def read(path):
    data = open(path, "rb").read()
    pos, pics, cur = 0, [], None

    def planes(p):
        out, q = [], pos
        for (pw, ph) in ((p["w"], p["h"]), (p["cw"], p["ch"]), (p["cw"], p["ch"])):
            out.append(np.frombuffer(data, np.int16, pw * ph, q).reshape(ph, pw).copy())
            q += 2 * pw * ph
        return out, q

    while pos < len(data):
        tag = data[pos:pos + 1]
        pos += 1
        if tag == b"P":
            hdr = np.frombuffer(data, np.int32, 16, pos)
            pos += 64
            cur = {"w": int(hdr[0]), "h": int(hdr[1]), "cw": int(hdr[2]), "ch": int(hdr[3]), "bd_luma": int(hdr[4]), "bd_chroma": int(hdr[5]),
                   "ctu": int(hdr[6]), "units_per_ctu": int(hdr[7]), "beta_offset_div2": int(hdr[8]), "tc_offset_div2": int(hdr[9]),
                   "cb_qp_offset": int(hdr[10]), "cr_qp_offset": int(hdr[11]), "n_ctus": int(hdr[12]), "ctus_x": int(hdr[13]), "poc": int(hdr[14])}
            cur["pre"], pos = planes(cur)
            uw, uh = (cur["w"] + 3) // 4, (cur["h"] + 3) // 4
            for k in ("bs_ver", "bs_hor", "nofilter"):
                cur[k] = np.zeros((uh, uw), np.uint8)
            cur["qp"] = np.zeros((uh, uw), np.int8)
            pics.append(cur)
        elif tag in (b"B", b"Q"):
            ctu, d = [int(v) for v in np.frombuffer(data, np.int32, 2, pos)]
            pos += 8
            n = cur["units_per_ctu"]
            ux0, uy0 = (ctu % cur["ctus_x"]) * n, (ctu // cur["ctus_x"]) * n
            uh, uw = cur["qp"].shape
            hh, ww = min(n, uh - uy0), min(n, uw - ux0)
            if tag == b"B":
                blk = np.frombuffer(data, np.uint8, n * n, pos).reshape(n, n)
                pos += n * n
                cur["bs_hor" if d else "bs_ver"][uy0:uy0 + hh, ux0:ux0 + ww] = blk[:hh, :ww]
            else:
                blk = np.frombuffer(data, np.uint8, 2 * n * n, pos).reshape(n, n, 2)
                pos += 2 * n * n
                cur["qp"][uy0:uy0 + hh, ux0:ux0 + ww] = blk[:hh, :ww, 0].view(np.int8)
                cur["nofilter"][uy0:uy0 + hh, ux0:ux0 + ww] = blk[:hh, :ww, 1]
        elif tag == b"E":
            cur["post"], pos = planes(cur)
        else:
            raise ValueError("bad record tag %r at %d" % (tag, pos - 1))
    return pics
