#!/usr/bin/env python3
"""Dev-time fixture generator (runs only where /root/reference exists).

Extracts the measured spectral DATA tables (CIE 1931 X/Y/Z, CIE D65, Canon EOS 100D
sensor curves, Al/Cu optical constants, BK7/SF11 glass IOR) from the numeric array
initialisers in the reference's src/color/spectra.cpp:10-499 and packs them into
quetzalcoatlus_b200/data/spectra_tables.bin, the file the host library loads
(quetzalcoatlus_b200/host/src/spectra.cpp).  Only numbers are taken -- no reference code.

Each literal is parsed the way the C++ compiler parses it (an unsuffixed literal is a
double which the array initialiser narrows to float; an f-suffixed one is a float), so the
packed float32 values are bit-identical to the reference's arrays.

File format (little endian): magic 'QZSPEC01', u32 n_tables, then per table:
char name[24] (NUL padded), u32 count, count x f32.
"""
import re
import struct
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/color/spectra.cpp")
OUT = Path(__file__).resolve().parent.parent / "quetzalcoatlus_b200" / "data" / "spectra_tables.bin"

NAMES = {
    "CIE_X": "CIE_X", "CIE_Y": "CIE_Y", "CIE_Z": "CIE_Z",
    "CIE_ILLUM_D6500": "D65",
    "CANON_EOS_100D_R": "CANON_R", "CANON_EOS_100D_G": "CANON_G", "CANON_EOS_100D_B": "CANON_B",
    "AL_IOR_VALUES": "AL_IOR", "AL_ABSORPTION_VALUES": "AL_ABSORPTION",
    "CU_IOR_VALUES": "CU_IOR", "CU_ABSORPTION_VALUES": "CU_ABSORPTION",
    "GLASS_BK7_VALUES": "GLASS_BK7_IOR", "GLASS_SF11_VALUES": "GLASS_SF11_IOR",
}


def parse_literal(tok: str) -> np.float32:
    if tok[-1] in "fF":
        return np.float32(tok[:-1])
    return np.float32(float(tok))  # double literal narrowed to float


def main() -> None:
    text = REF.read_text()
    text = re.sub(r"//[^\n]*", "", text)
    tables = []
    for m in re.finditer(r"(\w+)\s*(?:\[\s*\])?\s*=\s*\{([^{}]*)\}\s*;", text):
        ident, body = m.group(1), m.group(2)
        if ident not in NAMES:
            continue
        toks = [t.strip() for t in body.split(",") if t.strip()]
        vals = np.array([parse_literal(t) for t in toks], dtype=np.float32)
        tables.append((NAMES[ident], vals))
    missing = set(NAMES.values()) - {n for n, _ in tables}
    if missing:
        raise SystemExit(f"tables not found: {sorted(missing)}")
    with OUT.open("wb") as f:
        f.write(b"QZSPEC01")
        f.write(struct.pack("<I", len(tables)))
        for name, vals in tables:
            f.write(name.encode().ljust(24, b"\0"))
            f.write(struct.pack("<I", len(vals)))
            f.write(vals.astype("<f4").tobytes())
    for name, vals in tables:
        print(f"{name:16s} {len(vals):4d} floats  first={vals[0]:.6g} last={vals[-1]:.6g}")
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
