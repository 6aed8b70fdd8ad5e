"""FITS image I/O of the mapped images — mirror of src/shared/io.jl:12-58 (write_fits_image) and :106-141
(read_fits_image), the on-disk format on the far side of the deposit path.  Pure numpy: one
2880-byte-blocked header of 80-character cards + big-endian BITPIX=-64 data per image, each image of a
(N,N,N_images) map in its own HDU, with the reference's header keywords."""
from __future__ import annotations

import numpy as np

from .parameters import mappingParameters

_BLOCK = 2880


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        r = repr(float(value))
        if "e" in r or "E" in r:
            r = r.upper()
        elif r.endswith(".0"):
            r = r[:-1]
        v = f"{r:>20}"
    else:
        v = "'" + f"{str(value):<8}" + "'"
        v = f"{v:<20}"
    c = f"{key:<8}= {v}"
    if comment:
        c += f" / {comment}"
    return f"{c:<80}"[:80]


def _header_bytes(cards):
    txt = "".join(cards) + f"{'END':<80}"
    pad = (-len(txt)) % _BLOCK
    return (txt + " " * pad).encode("ascii")


def write_fits_image(filename, image, par: mappingParameters, units="[i.u.]", snap=0):
    """Writes a mapped image to a FITS file and stores the essential mapping parameters in the header."""
    image = np.asarray(image, dtype=np.float64)
    if image.ndim == 2:
        image = image[:, :, None]
    z_slice = abs(par.z_lim[0]) + abs(par.z_lim[1])
    keys = [("SNAP", int(snap), "snapshot number"),
            ("CENTER_X", float(par.center[0]), "image center (x)"), ("CENTER_Y", float(par.center[1]), "image center (y)"),
            ("CENTER_Z", float(par.center[2]), "image center (z)"),
            ("XMIN", float(par.x_lim[0]), "x limit left"), ("XMAX", float(par.x_lim[1]), "x limit right"),
            ("YMIN", float(par.y_lim[0]), "y limit left"), ("YMAX", float(par.y_lim[1]), "y limit right"),
            ("ZMIN", float(par.z_lim[0]), "z limit left"), ("ZMAX", float(par.z_lim[1]), "z limit right"),
            ("BOXSIZE", float(par.boxsize), "size of the image"), ("Z_SLICE", float(z_slice), "depth of the integrated slice"),
            ("NPIXELS", int(par.Npixels.max()), "image resolution"), ("PIX_SIZE", float(par.pixelSideLength), "pixel size"),
            ("UNITS", units, "units of the image")]
    with open(filename, "wb") as f:
        for n in range(image.shape[2]):
            img = image[:, :, n]
            n1, n2 = img.shape  # Julia (n1, n2) column-major: NAXIS1 = n1 is the fast axis
            if n == 0:
                cards = [_card("SIMPLE", True, "file does conform to FITS standard")]
            else:
                cards = [f"{'XTENSION= ' + repr('IMAGE   '):<30} / IMAGE extension"[:80].ljust(80)]
            cards += [_card("BITPIX", -64, "number of bits per data pixel"), _card("NAXIS", 2, "number of data axes"),
                      _card("NAXIS1", n1, "length of data axis 1"), _card("NAXIS2", n2, "length of data axis 2")]
            cards += [_card("EXTEND", True, "FITS dataset may contain extensions")] if n == 0 else \
                     [_card("PCOUNT", 0, "required keyword; must = 0"), _card("GCOUNT", 1, "required keyword; must = 1")]
            cards += [_card(k, v, c) for k, v, c in keys]
            f.write(_header_bytes(cards))
            data = np.asfortranarray(img).astype(">f8").tobytes(order="F")
            f.write(data + b"\0" * ((-len(data)) % _BLOCK))


def _parse_value(raw):
    raw = raw.split("/")[0].strip() if not raw.strip().startswith("'") else raw
    s = raw.strip()
    if s.startswith("'"):
        return s[1:s.index("'", 1)].rstrip()
    if s in ("T", "F"):
        return s == "T"
    try:
        return int(s)
    except ValueError:
        return float(s.replace("D", "E"))


def read_fits_hdus(filename):
    """All image HDUs of a FITS file as (header dict, array indexed like the Julia array [axis1, axis2])."""
    raw = open(filename, "rb").read()
    pos, out = 0, []
    while pos < len(raw):
        hdr = {}
        while True:
            block = raw[pos:pos + _BLOCK].decode("latin1")
            pos += _BLOCK
            done = False
            for i in range(0, _BLOCK, 80):
                card = block[i:i + 80]
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if card[8:10] == "= ":
                    hdr[key] = _parse_value(card[10:])
            if done:
                break
        naxis = int(hdr.get("NAXIS", 0))
        shape = [int(hdr[f"NAXIS{i + 1}"]) for i in range(naxis)]
        nbytes = abs(int(hdr.get("BITPIX", 8))) // 8 * int(np.prod(shape)) if naxis else 0
        if naxis:
            dt = {-64: ">f8", -32: ">f4", 16: ">i2", 32: ">i4", 64: ">i8", 8: "u1"}[int(hdr["BITPIX"])]
            arr = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape)), offset=pos).astype(dt[1:])
            out.append((hdr, arr.reshape(shape, order="F")))
        pos += nbytes + ((-nbytes) % _BLOCK)
    return out


def read_fits_image(filename, Nimage=1, verbose=False):
    """Returns `(image, par, snap, units)` like the reference's read_fits_image."""
    hdr, image = read_fits_hdus(filename)[Nimage - 1]
    par = mappingParameters(x_lim=[hdr["XMIN"], hdr["XMAX"]], y_lim=[hdr["YMIN"], hdr["YMAX"]],
                            z_lim=[hdr["ZMIN"], hdr["ZMAX"]], Npixels=int(hdr["NPIXELS"]), boxsize=hdr["BOXSIZE"])
    return image, par, hdr["SNAP"], hdr["UNITS"]
