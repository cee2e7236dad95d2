"""Generates the six cube-map faces of the synthetic sky used by the cubemap-environment EXTENSION (the reference ships no
cubemap asset: SURVEY.md F3).  Closed-form sky (elevation gradient, sun disc with halo, soft sine bands, brown ground), no RNG;
24-bit BMPs, 128x128, written so that decoded row 0 is the top row.  Face / uv convention: oracle/orc_scene.hpp Environment.
usage: python scenes/skybox/make_skybox.py"""
import math
import os
import struct

N = 128
HERE = os.path.dirname(os.path.abspath(__file__))
SUN = (-0.45, 0.55, 0.70)


def sky(d):
    l = math.sqrt(sum(c * c for c in d))
    x, y, z = (c / l for c in d)
    if y < 0:   # ground: darker towards the nadir, a faint grid of sine bands
        g = 0.25 + 0.2 * (1 + y)
        band = 0.04 * math.sin(9 * x / (0.2 - y)) * math.sin(9 * z / (0.2 - y))
        return (g * 0.9 + band, g * 0.7 + band, g * 0.45 + band)
    t = y ** 0.5
    r, g, b = 0.85 - 0.55 * t, 0.9 - 0.4 * t, 1.0 - 0.1 * t   # pale horizon -> blue zenith
    c = 0.06 * math.sin(7 * x + 3 * z) * math.sin(5 * z - 2 * x) * (1 - t)   # soft bands near the horizon
    sl = math.sqrt(sum(s * s for s in SUN))
    cs = (x * SUN[0] + y * SUN[1] + z * SUN[2]) / sl
    halo = max(0.0, cs) ** 64 * 0.6 + (1.0 if cs > 0.9985 else 0.0)
    return (r + c + halo, g + c + halo * 0.95, b + c + halo * 0.8)


def direction(face, sx, sy):
    return {"posx": (1, -sy, -sx), "negx": (-1, -sy, sx), "posy": (sx, 1, sy), "negy": (sx, -1, -sy),
            "posz": (sx, -sy, 1), "negz": (-sx, -sy, -1)}[face]


def write_bmp(path, rows):   # rows[0] = top row, pixels (r, g, b) in 0..255
    h, w = len(rows), len(rows[0])
    stride = (w * 3 + 3) // 4 * 4
    data = bytearray()
    for row in reversed(rows):   # BMP stores the bottom row first
        line = bytearray()
        for (r, g, b) in row:
            line += bytes((b, g, r))
        line += bytes(stride - len(line))
        data += line
    hdr = struct.pack("<2sIHHI", b"BM", 54 + len(data), 0, 0, 54)
    info = struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(data), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + info + data)


def main():
    for face in ("posx", "negx", "posy", "negy", "posz", "negz"):
        rows = []
        for ty in range(N):
            row = []
            for tx in range(N):
                c = sky(direction(face, 2 * tx / (N - 1) - 1, 2 * ty / (N - 1) - 1))
                row.append(tuple(max(0, min(255, int(round(255 * v ** (1 / 2.2))))) for v in (max(0.0, min(1.0, q)) for q in c)))
            rows.append(row)
        write_bmp(os.path.join(HERE, face + ".bmp"), rows)


if __name__ == "__main__":
    main()
