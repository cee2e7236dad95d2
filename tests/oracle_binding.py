"""ctypes binding of the CPU oracle (oracle/liborc.so, oracle/liborc_count.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs load this.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OrcStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("flops", C.c_uint64),
                ("prepass_rays", C.c_uint64), ("prepass_flops", C.c_uint64), ("csg_max_crossings", C.c_uint64),
                ("seconds", C.c_double), ("gi_bounce_rays", C.c_uint64)]


def _bind(path):
    lib = C.CDLL(path)
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_scene_load.restype = C.c_void_p
    lib.orc_scene_load.argtypes = [C.c_char_p]
    lib.orc_scene_free.argtypes = [C.c_void_p]
    lib.orc_scene_free.restype = None
    lib.orc_scene_set_frame_size.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.orc_scene_set_frame_size.restype = None
    lib.orc_scene_get_frame_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.orc_scene_get_frame_size.restype = None
    lib.orc_scene_override.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_scene_override.restype = None
    lib.orc_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    lib.orc_scene_info.restype = None
    lib.orc_camera_vectors.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.orc_camera_vectors.restype = None
    lib.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_uint64, C.POINTER(OrcStats)]
    lib.orc_render_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint, C.c_int, C.c_uint64,
                                    C.POINTER(OrcStats)]
    lib.orc_render_pixel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_float),
                                     C.POINTER(C.c_double)]
    lib.orc_environment.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.orc_pixel_diag.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_double, C.POINTER(C.c_double)]
    lib.orc_pack_rgb32.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.orc_pack_rgb32.restype = None
    lib.orc_srgb_lut.argtypes = [C.c_void_p]
    lib.orc_srgb_lut.restype = None
    lib.orc_kat_intersect.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.c_double, C.POINTER(C.c_double)]
    lib.orc_kat_checker.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_float)]
    lib.orc_kat_checker.restype = None
    lib.orc_kat_shell_sort.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    lib.orc_kat_shell_sort.restype = None
    lib.orc_kat_decode_bmp.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p,
                                       C.c_size_t]
    lib.orc_texture_texels.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p,
                                       C.c_size_t]
    lib.orc_rng_u31.argtypes = [C.c_uint64] + [C.c_uint32] * 5
    lib.orc_rng_u31.restype = C.c_uint32
    return lib


_libs = {}


def oracle_lib(count_flops=False):
    name = "liborc_count.so" if count_flops else "liborc.so"
    if name not in _libs:
        path = os.path.join(ROOT, "oracle", name)
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _libs[name] = _bind(path)
    return _libs[name]


class OracleScene:
    RNG_LIBC, RNG_PINNED = 0, 1

    def __init__(self, path, count_flops=False):
        self.lib = oracle_lib(count_flops)
        self._h = self.lib.orc_scene_load(os.fspath(path).encode())
        if not self._h:
            raise RuntimeError(self.lib.orc_last_error().decode())

    def close(self):
        if self._h:
            self.lib.orc_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_frame_size(self, w, h):
        self.lib.orc_scene_set_frame_size(self._h, w, h)

    @property
    def frame_size(self):
        w, h = C.c_uint32(), C.c_uint32()
        self.lib.orc_scene_get_frame_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def override(self, aa=-1, dof=-1, prepass=-1, num_samples=-1):
        self.lib.orc_scene_override(self._h, int(aa), int(dof), int(prepass), int(num_samples))

    def info(self):
        out = (C.c_int32 * 8)()
        self.lib.orc_scene_info(self._h, out)
        keys = ["nodes", "geometries", "shaders", "textures", "lights", "aa", "dof", "num_samples"]
        return dict(zip(keys, list(out)))

    def camera_vectors(self):
        out = (C.c_double * 21)()
        self.lib.orc_camera_vectors(self._h, out)
        return np.array(list(out)).reshape(7, 3)

    def render(self, threads=0, rng_mode=1, seed=0):
        w, h = self.frame_size
        rgb = np.zeros((h, w, 3), np.float32)
        st = OrcStats()
        if self.lib.orc_render(self._h, rgb.ctypes.data, threads, rng_mode, seed, C.byref(st)) != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return rgb, st

    def render_rows(self, y0, y1, threads=0, rng_mode=1, seed=0):
        w, h = self.frame_size
        rgb = np.zeros((y1 - y0, w, 3), np.float32)
        st = OrcStats()
        if self.lib.orc_render_rows(self._h, rgb.ctypes.data, y0, y1, threads, rng_mode, seed, C.byref(st)) != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return rgb, st

    def render_pixel(self, x, y, rng_mode=1, seed=0):
        rgb = (C.c_float * 3)()
        hit = (C.c_double * 10)()
        if self.lib.orc_render_pixel(self._h, x, y, rng_mode, seed, rgb, hit) != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return np.array(list(rgb), np.float32), np.array(list(hit))

    def environment(self, direction):
        """Environment.getEnvironment(dir) -> (is_cubemap, rgb)."""
        d = (C.c_double * 3)(*direction)
        rgb = (C.c_float * 3)()
        kind = self.lib.orc_environment(self._h, d, rgb)
        return bool(kind), np.array(list(rgb), np.float32)

    def pixel_diag(self, x, y, rng_mode=1, seed=0, eps=1e-7):
        """Conditioning of one pixel of the oracle's own image -> dict(max_dist, min_gap, shift_delta, rgb)."""
        out = (C.c_double * 6)()
        if self.lib.orc_pixel_diag(self._h, x, y, rng_mode, seed, eps, out) != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return {"max_dist": out[0], "min_gap": out[1], "shift_delta": out[2], "rgb": (out[3], out[4], out[5])}

    def texture_texels(self, idx):
        w, h = C.c_uint32(), C.c_uint32()
        rc = self.lib.orc_texture_texels(self._h, idx, C.byref(w), C.byref(h), None, 0)
        if rc != 0:
            return None
        out = np.zeros((h.value, w.value, 3), np.float32)
        self.lib.orc_texture_texels(self._h, idx, C.byref(w), C.byref(h), out.ctypes.data, out.size)
        return out


def pack_rgb32(rgb):
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.zeros(rgb.shape[:-1], np.uint32)
    oracle_lib().orc_pack_rgb32(rgb.ctypes.data, out.size, out.ctypes.data)
    return out


def srgb_lut():
    out = np.zeros(4097, np.uint8)
    oracle_lib().orc_srgb_lut(out.ctypes.data)
    return out


def parity_report(gpu_rgb, ref_rgb, gpu_argb=None):
    """The north-star bar: per-pixel float RGB within 1e-3 abs; <= 0.1 % of 8-bit pixels off by > 1 LSB."""
    d = np.abs(gpu_rgb.astype(np.float64) - ref_rgb.astype(np.float64))
    per_px = d.max(axis=-1)
    a = pack_rgb32(gpu_rgb) if gpu_argb is None else gpu_argb
    b = pack_rgb32(ref_rgb)
    lsb = np.zeros(a.shape, np.int32)
    for sh in (0, 8, 16):
        lsb = np.maximum(lsb, np.abs(((a >> sh) & 255).astype(np.int32) - ((b >> sh) & 255).astype(np.int32)))
    n = per_px.size
    return {
        "max_abs": float(per_px.max()),
        "px_over_1e-3": int((per_px > 1e-3).sum()),
        "frac_over_1e-3": float((per_px > 1e-3).sum() / n),
        "px_over_1lsb": int((lsb > 1).sum()),
        "frac_over_1lsb": float((lsb > 1).sum() / n),
        "n": int(n),
    }
