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


def write_fits_image(filename, image, par: mappingParameters = None, units="[i.u.]", snap=0):
    """Writes a mapped image to a FITS file and stores the essential mapping parameters in the header
    (io.jl:12-58).  Without `par` this is the reference's second method (io.jl:62-89): an allsky / any-dimensional
    image in the primary HDU with only SNAP and UNITS in the header."""
    if par is None:
        return _write_plain_fits(filename, image, units, snap)
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


def _write_plain_fits(filename, image, units, snap):
    image = np.asarray(image, dtype=np.float64)
    cards = [_card("SIMPLE", True, "file does conform to FITS standard"),
             _card("BITPIX", -64, "number of bits per data pixel"), _card("NAXIS", image.ndim, "number of data axes")]
    cards += [_card(f"NAXIS{i + 1}", n, f"length of data axis {i + 1}") for i, n in enumerate(image.shape)]
    cards += [_card("EXTEND", True, "FITS dataset may contain extensions"),
              _card("SNAP", int(snap), "snapshot number"), _card("UNITS", units, "units of the image")]
    with open(filename, "wb") as f:
        f.write(_header_bytes(cards))
        data = np.asfortranarray(image).astype(">f8").tobytes(order="F")
        f.write(data + b"\0" * ((-len(data)) % _BLOCK))


def read_allsky_fits_image(filename, verbose=False):
    """Returns `(image, snap, units)` (io.jl:146-186)."""
    hdr, image = read_fits_hdus(filename)[0]
    return image, hdr["SNAP"], hdr["UNITS"]


def save_healpix_fits(filename, pixels, ordering="RING", unit="", extname="MAP", overwrite=True):
    """`saveToFITS(map, filename)` of Healpix.jl, as called by distributed_allsky_map
    (src/distributed_mapping/healpix.jl:73-77, which removes an existing file first): an empty primary HDU and one
    BINTABLE extension with a single Float64 column PIXVALS, one pixel per row, and the HEALPix keywords PIXTYPE,
    ORDERING, NSIDE, FIRSTPIX, LASTPIX, INDXSCHM, OBJECT.  Healpix.jl is not part of the reference tree: the layout
    follows the HEALPix FITS convention it implements (readable by healpy / Healpix.jl `readMapFromFITS`); the exact
    card order and comments are unpinned."""
    import os
    pixels = np.ascontiguousarray(pixels, dtype=np.float64).ravel()
    npix = pixels.shape[0]
    nside = int(round((npix / 12) ** 0.5))
    if 12 * nside * nside != npix:
        raise ValueError(f"{npix} is not a valid HEALPix pixel count")
    if os.path.isfile(filename):
        if not overwrite:
            raise FileExistsError(filename)
        os.remove(filename)
    primary = [_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", 8, "number of bits per data pixel"),
               _card("NAXIS", 0, "number of data axes"), _card("EXTEND", True, "FITS dataset may contain extensions")]
    ext = [f"{'XTENSION= ' + repr('BINTABLE'):<30} / binary table extension"[:80].ljust(80),
           _card("BITPIX", 8, "8-bit bytes"), _card("NAXIS", 2, "2-dimensional binary table"),
           _card("NAXIS1", 8, "width of table in bytes"), _card("NAXIS2", npix, "number of rows in table"),
           _card("PCOUNT", 0, "size of special data area"), _card("GCOUNT", 1, "one data group (required keyword)"),
           _card("TFIELDS", 1, "number of fields in each row"), _card("TTYPE1", "PIXVALS", "label for field   1"),
           _card("TFORM1", "1D", "data format of field: 8-byte DOUBLE")]
    if unit:
        ext.append(_card("TUNIT1", unit, "physical unit of field"))
    ext += [_card("EXTNAME", extname, "name of this binary table extension"),
            _card("PIXTYPE", "HEALPIX", "HEALPIX pixelization"), _card("ORDERING", ordering, "Pixel ordering scheme"),
            _card("NSIDE", nside, "Value of NSIDE"), _card("FIRSTPIX", 0, "First pixel (0 based)"),
            _card("LASTPIX", npix - 1, "Last pixel (0 based)"), _card("INDXSCHM", "IMPLICIT", "Indexing: IMPLICIT or EXPLICIT"),
            _card("OBJECT", "FULLSKY", "Sky coverage, either FULLSKY or PARTIAL")]
    with open(filename, "wb") as f:
        f.write(_header_bytes(primary))
        f.write(_header_bytes(ext))
        data = pixels.astype(">f8").tobytes()
        f.write(data + b"\0" * ((-len(data)) % _BLOCK))


def read_healpix_fits(filename):
    """Reads back a single-column HEALPix FITS table (any repeat count of D/E per row): `(pixels, header)`."""
    raw = open(filename, "rb").read()
    pos = 0
    while pos < len(raw):
        hdr, pos = _read_header(raw, pos)
        naxis = int(hdr.get("NAXIS", 0))
        nbytes = abs(int(hdr.get("BITPIX", 8))) // 8 * int(np.prod([hdr[f"NAXIS{i + 1}"] for i in range(naxis)])) \
            if naxis else 0
        if hdr.get("XTENSION", "").strip() == "BINTABLE":
            form = str(hdr["TFORM1"]).strip()
            rep = int(form[:-1]) if form[:-1] else 1
            dt = {"D": ">f8", "E": ">f4"}[form[-1]]
            if int(hdr["TFIELDS"]) != 1 or int(hdr["NAXIS1"]) != rep * int(dt[2]):
                raise ValueError("only single-column HEALPix tables are supported")
            px = np.frombuffer(raw, dtype=dt, count=rep * int(hdr["NAXIS2"]), offset=pos).astype(np.float64)
            return px, hdr
        pos += nbytes + ((-nbytes) % _BLOCK)
    raise ValueError("no BINTABLE extension found")


def get_map_grid_3D(par: mappingParameters):
    """Pixel-centre coordinates of the 3D grid (src/shared/reconstruct_grid.jl:27-46; the z axis starts from
    `y_lim[1]` there — reproduced)."""
    i = np.arange(1, int(par.Npixels[0]) + 1)
    x = par.x_lim[0] + (i - 0.5) * par.pixelSideLength
    y = par.y_lim[0] + (np.arange(1, int(par.Npixels[1]) + 1) - 0.5) * par.pixelSideLength
    z = par.y_lim[0] + (np.arange(1, int(par.Npixels[2]) + 1) - 0.5) * par.pixelSideLength
    return x, y, z


def write_vtk_image(filename, image, image_name, par: mappingParameters, units="[i.u.]", snap=0):
    """`write_vtk_image` (src/shared/vtk.jl:10-22): WriteVTK's `vtk_grid(filename, x, y, z)` with three coordinate
    vectors writes a RECTILINEAR grid (`.vtr`), the map as point data `image_name` and `Units` / `Snap` as field
    data.  Written here as VTK XML with raw appended binary data (little endian, UInt64 headers); returns the file
    name with the extension WriteVTK would add."""
    import struct
    image = np.asarray(image, dtype=np.float64)
    x, y, z = get_map_grid_3D(par)
    if image.shape != (x.size, y.size, z.size):
        raise ValueError(f"image has shape {image.shape}, the grid is {(x.size, y.size, z.size)}")
    if not filename.endswith(".vtr"):
        filename += ".vtr"
    ustr = units.encode("utf-8") + b"\0"
    blocks, offsets, off = [], [], 0
    for payload in (image.tobytes(order="F"), np.frombuffer(ustr, dtype=np.uint8).tobytes(),
                    np.array([int(snap)], dtype=np.int64).tobytes(), x.tobytes(), y.tobytes(), z.tobytes()):
        offsets.append(off)
        blocks.append(struct.pack("<Q", len(payload)) + payload)
        off += 8 + len(payload)
    ext = f"0 {x.size - 1} 0 {y.size - 1} 0 {z.size - 1}"
    xml = (f'<?xml version="1.0" encoding="utf-8"?>\n'
           f'<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
           f'  <RectilinearGrid WholeExtent="{ext}">\n'
           f'    <FieldData>\n'
           f'      <Array type="String" Name="Units" NumberOfTuples="1" format="appended" offset="{offsets[1]}"/>\n'
           f'      <DataArray type="Int64" Name="Snap" NumberOfTuples="1" format="appended" offset="{offsets[2]}"/>\n'
           f'    </FieldData>\n'
           f'    <Piece Extent="{ext}">\n'
           f'      <PointData>\n'
           f'        <DataArray type="Float64" Name="{image_name}" NumberOfComponents="1" format="appended" '
           f'offset="{offsets[0]}"/>\n'
           f'      </PointData>\n'
           f'      <CellData/>\n'
           f'      <Coordinates>\n'
           f'        <DataArray type="Float64" Name="x" NumberOfComponents="1" format="appended" offset="{offsets[3]}"/>\n'
           f'        <DataArray type="Float64" Name="y" NumberOfComponents="1" format="appended" offset="{offsets[4]}"/>\n'
           f'        <DataArray type="Float64" Name="z" NumberOfComponents="1" format="appended" offset="{offsets[5]}"/>\n'
           f'      </Coordinates>\n'
           f'    </Piece>\n'
           f'  </RectilinearGrid>\n'
           f'  <AppendedData encoding="raw">\n_')
    with open(filename, "wb") as f:
        f.write(xml.encode("utf-8"))
        for b in blocks:
            f.write(b)
        f.write(b"\n  </AppendedData>\n</VTKFile>\n")
    return filename


def _read_header(raw, pos):
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
            return hdr, pos


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
