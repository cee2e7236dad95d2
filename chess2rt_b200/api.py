"""ctypes bindings of include/c2rt.h (libc2rt.so) and of the host mirror (libc2rt_host.so)."""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))


class C2rtError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"c2rt status {status}: {message}")
        self.status = status


def _load(name):
    path = os.path.join(os.environ.get("C2RT_LIB_DIR") or _PKG, name)  # C2RT_LIB_DIR: tuning variants (build.build_variant)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m chess2rt_b200.build` (or __graft_entry__.build()). "
            "There is no Python/CPU fallback for the render path.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


lib = _load("libc2rt.so")
host_lib = _load("libc2rt_host.so")


class Camera(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("up_left", C.c_double * 3), ("up_right", C.c_double * 3),
                ("down_left", C.c_double * 3), ("right_dir", C.c_double * 3), ("up_dir", C.c_double * 3),
                ("front_dir", C.c_double * 3), ("frame_width", C.c_uint32), ("frame_height", C.c_uint32),
                ("dof", C.c_int32), ("num_samples", C.c_uint32), ("focal_plane_dist", C.c_double),
                ("disc_multiplier", C.c_double), ("stereo_separation", C.c_double)]


class Settings(C.Structure):
    _fields_ = [("frame_width", C.c_uint32), ("frame_height", C.c_uint32), ("aa_enabled", C.c_int32),
                ("gi_enabled", C.c_int32), ("prepass_enabled", C.c_int32), ("prepass_only", C.c_int32),
                ("max_trace_depth", C.c_uint32), ("ambient_light", C.c_float * 3), ("rng_seed", C.c_uint64),
                ("count_rays", C.c_int32), ("bucket_size", C.c_uint32),
                ("paths_per_pixel", C.c_uint32), ("reserved", C.c_uint32)]


class Band(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("n_ranks", C.c_uint32), ("band_rows", C.c_uint32), ("compact", C.c_uint32),
                ("done_flags", C.c_void_p), ("frame_no", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("primary_rays", C.c_uint64),
                ("shadow_rays", C.c_uint64), ("n_gpus", C.c_uint32), ("launches", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("node", C.c_int32), ("reserved", C.c_int32), ("dist", C.c_double), ("p", C.c_double * 3),
                ("normal", C.c_double * 3), ("u", C.c_double), ("v", C.c_double)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("n_nodes", C.c_uint32), ("node_geom", C.POINTER(C.c_int32)), ("node_shader", C.POINTER(C.c_int32)),
        ("node_transform", C.POINTER(C.c_double)), ("node_inverse", C.POINTER(C.c_double)),
        ("node_inverse_t", C.POINTER(C.c_double)), ("node_offset", C.POINTER(C.c_double)),
        ("n_geoms", C.c_uint32), ("geom_type", C.POINTER(C.c_int32)), ("geom_params", C.POINTER(C.c_double)),
        ("geom_left", C.POINTER(C.c_int32)), ("geom_right", C.POINTER(C.c_int32)),
        ("n_shaders", C.c_uint32), ("shader_type", C.POINTER(C.c_int32)), ("shader_color", C.POINTER(C.c_float)),
        ("shader_texture", C.POINTER(C.c_int32)), ("shader_exponent", C.POINTER(C.c_double)),
        ("shader_strength", C.POINTER(C.c_float)),
        ("n_textures", C.c_uint32), ("tex_type", C.POINTER(C.c_int32)), ("tex_colors", C.POINTER(C.c_float)),
        ("tex_params", C.POINTER(C.c_double)), ("tex_width", C.POINTER(C.c_int32)), ("tex_height", C.POINTER(C.c_int32)),
        ("tex_texel_offset", C.POINTER(C.c_uint64)), ("texels", C.POINTER(C.c_float)), ("n_texels", C.c_uint64),
        ("n_lights", C.c_uint32), ("light_pos", C.POINTER(C.c_double)), ("light_color", C.POINTER(C.c_float)),
        ("light_power", C.POINTER(C.c_float)),
        ("env_type", C.c_int32), ("env_reserved", C.c_int32), ("env_face_width", C.c_int32 * 6), ("env_face_height", C.c_int32 * 6),
        ("env_face_texel_offset", C.c_uint64 * 6),
    ]


# every symbol include/c2rt.h declares (tests/test_abi.py checks the library exports them all)
C_ABI_SYMBOLS = [
    "c2rt_init", "c2rt_shutdown", "c2rt_abi_version", "c2rt_device_count", "c2rt_last_error",
    "c2rt_scene_create", "c2rt_scene_destroy", "c2rt_render", "c2rt_cancel", "c2rt_render_device", "c2rt_read_ray_counters",
    "c2rt_deinterleave", "c2rt_render_pixel", "c2rt_band_rows_owned", "c2rt_rng_u31", "c2rt_srgb_table",
    "c2rt_frame_alloc", "c2rt_frame_free", "c2rt_frame_export", "c2rt_frame_import", "c2rt_frame_unimport", "c2rt_frame_memset", "c2rt_frame_download", "c2rt_pin_host_buffer", "c2rt_unpin_host_buffer", "c2rt_gate",
    "c2rt_measure_fma_peak", "c2rt_selftest_device_pool",
]

lib.c2rt_last_error.restype = C.c_char_p
lib.c2rt_selftest_device_pool.argtypes = [C.c_int, C.c_int]
lib.c2rt_selftest_device_pool.restype = C.c_longlong
lib.c2rt_init.argtypes = [C.c_int, C.POINTER(C.c_int)]
lib.c2rt_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
lib.c2rt_scene_destroy.argtypes = [C.c_void_p]
lib.c2rt_scene_destroy.restype = None
lib.c2rt_render.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
lib.c2rt_render_device.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.POINTER(Band), C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.POINTER(Stats)]
lib.c2rt_read_ray_counters.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.c2rt_deinterleave.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_void_p]
lib.c2rt_render_pixel.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Settings), C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(Hit)]
lib.c2rt_band_rows_owned.argtypes = [C.c_uint32] * 4
lib.c2rt_band_rows_owned.restype = C.c_uint32
lib.c2rt_rng_u31.argtypes = [C.c_uint64] + [C.c_uint32] * 5
lib.c2rt_rng_u31.restype = C.c_uint32
lib.c2rt_srgb_table.argtypes = [C.c_void_p]
lib.c2rt_srgb_table.restype = None
lib.c2rt_frame_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
lib.c2rt_frame_free.argtypes = [C.c_void_p]
lib.c2rt_frame_export.argtypes = [C.c_void_p, C.c_void_p]
lib.c2rt_frame_import.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
lib.c2rt_frame_unimport.argtypes = [C.c_void_p]
lib.c2rt_frame_memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
lib.c2rt_frame_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.c2rt_pin_host_buffer.argtypes = [C.c_void_p, C.c_size_t]
lib.c2rt_unpin_host_buffer.argtypes = [C.c_void_p]
lib.c2rt_gate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
lib.c2rt_measure_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]

host_lib.c2rt_host_last_error.restype = C.c_char_p
host_lib.c2rt_host_scene_load.argtypes = [C.c_char_p]
host_lib.c2rt_host_scene_load.restype = C.c_void_p
host_lib.c2rt_host_scene_free.argtypes = [C.c_void_p]
host_lib.c2rt_host_scene_free.restype = None
host_lib.c2rt_host_scene_set_frame_size.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
host_lib.c2rt_host_scene_set_frame_size.restype = None
host_lib.c2rt_host_scene_get_frame_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
host_lib.c2rt_host_scene_get_frame_size.restype = None
host_lib.c2rt_host_scene_override.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
host_lib.c2rt_host_scene_override.restype = None
host_lib.c2rt_host_scene_desc.argtypes = [C.c_void_p]
host_lib.c2rt_host_scene_desc.restype = C.POINTER(SceneDesc)
host_lib.c2rt_host_frame_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Camera), C.POINTER(Settings)]
host_lib.c2rt_host_frame_blocks.restype = None
host_lib.c2rt_host_device_scene.argtypes = [C.c_void_p]
host_lib.c2rt_host_device_scene.restype = C.c_void_p
host_lib.c2rt_host_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Stats)]
host_lib.c2rt_host_render_pixel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(Hit)]
host_lib.c2rt_host_camera_rotate.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
host_lib.c2rt_host_camera_rotate.restype = None
host_lib.c2rt_host_camera_move.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
host_lib.c2rt_host_camera_move.restype = None
host_lib.c2rt_host_save_bmp.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
host_lib.c2rt_host_save_bmp.restype = C.c_size_t
host_lib.c2rt_host_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
host_lib.c2rt_host_scene_info.restype = None
host_lib.c2rt_host_decode_bmp.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p,
                                          C.c_size_t]


def _check(rc):
    if rc != 0:
        raise C2rtError(rc, lib.c2rt_last_error().decode())


def _check_host(rc):
    if rc != 0:
        raise C2rtError(rc, host_lib.c2rt_host_last_error().decode())


def device_count():
    return lib.c2rt_device_count()


def init(n_gpus=1, device_ids=None):
    ids = (C.c_int * n_gpus)(*device_ids) if device_ids is not None else None
    _check(lib.c2rt_init(n_gpus, ids))


def shutdown():
    lib.c2rt_shutdown()


def band_rows_owned(height, rank, n_ranks, band_rows):
    return lib.c2rt_band_rows_owned(height, rank, n_ranks, band_rows)


def rng_u31(seed, px, py, tap, sample, draw):
    return lib.c2rt_rng_u31(seed, px, py, tap, sample, draw)


def srgb_table():
    out = np.zeros(4097, np.uint8)
    lib.c2rt_srgb_table(out.ctypes.data)
    return out


def measure_fma_peak(fp64=False):
    tf, mhz = C.c_double(), C.c_double()
    _check(lib.c2rt_measure_fma_peak(int(fp64), C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


class HostScene:
    """A scene file loaded by the reference-compatible host loader (rt::parseSceneFromFile)."""

    def __init__(self, path):
        self._h = host_lib.c2rt_host_scene_load(os.fspath(path).encode())
        if not self._h:
            raise C2rtError(-1, host_lib.c2rt_host_last_error().decode())
        self.path = os.fspath(path)

    def close(self):
        if self._h:
            host_lib.c2rt_host_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_frame_size(self, w, h):
        host_lib.c2rt_host_scene_set_frame_size(self._h, w, h)

    @property
    def frame_size(self):
        w, h = C.c_uint32(), C.c_uint32()
        host_lib.c2rt_host_scene_get_frame_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def override(self, aa=-1, dof=-1, prepass=-1, num_samples=-1):
        host_lib.c2rt_host_scene_override(self._h, int(aa), int(dof), int(prepass), int(num_samples))

    def info(self):
        out = (C.c_int32 * 8)()
        host_lib.c2rt_host_scene_info(self._h, out)
        keys = ["nodes", "geometries", "shaders", "textures", "lights", "aa", "dof", "num_samples"]
        return dict(zip(keys, list(out)))

    def desc(self):
        p = host_lib.c2rt_host_scene_desc(self._h)
        if not p:
            raise C2rtError(-1, host_lib.c2rt_host_last_error().decode())
        return p

    def frame_blocks(self, seed=0, count_rays=False):
        cam, st = Camera(), Settings()
        host_lib.c2rt_host_frame_blocks(self._h, seed, int(count_rays), C.byref(cam), C.byref(st))
        return cam, st

    def device_scene(self):
        p = host_lib.c2rt_host_device_scene(self._h)
        if not p:
            raise C2rtError(-3, host_lib.c2rt_host_last_error().decode())
        return p

    def render(self, argb=False, seed=0, count_rays=False, out=None, out_argb=None, argb_only=False):
        """Renderer(scene, output).renderRT() with HOST buffers -> (rgb[H,W,3] float32, argb[H,W] uint32|None, Stats).
        argb_only: no float frame is written or copied (rgb is None in the result)."""
        w, h = self.frame_size
        rgb = None if argb_only else (out if out is not None else np.empty((h, w, 3), np.float32))
        a = out_argb if out_argb is not None else (np.empty((h, w), np.uint32) if (argb or argb_only) else None)
        st = Stats()
        rc = host_lib.c2rt_host_render(self._h, rgb.ctypes.data if rgb is not None else None, a.ctypes.data if a is not None else None,
                                       seed, int(count_rays), C.byref(st))
        self.cancelled = rc == 1   # C2RT_CANCELLED: a c2rt_cancel reached the frame (partial image), not an error
        if rc != 1:
            _check_host(rc)
        return rgb, a, st

    def camera_rotate(self, d_yaw, d_roll=0.0, d_pitch=0.0):
        host_lib.c2rt_host_camera_rotate(self._h, d_yaw, d_roll, d_pitch)

    def camera_move(self, dx, dy, dz):
        host_lib.c2rt_host_camera_move(self._h, dx, dy, dz)

    def render_pixel(self, x, y):
        rgb = (C.c_float * 3)()
        hit = Hit()
        _check_host(host_lib.c2rt_host_render_pixel(self._h, x, y, rgb, C.byref(hit)))
        return np.array(list(rgb), np.float32), hit


def save_bmp(argb, pad_rows=False):
    """saveBmp (imageio/bmp.d:195-237) of a packed uint32 plane [H, W] -> bytes."""
    a = np.ascontiguousarray(argb, np.uint32)
    h, w = a.shape
    out = np.zeros(54 + (w * 3 + 3) // 4 * 4 * h, np.uint8)
    n = host_lib.c2rt_host_save_bmp(a.ctypes.data, w, h, int(pad_rows), out.ctypes.data, out.size)
    return out[:n].tobytes()


def cancel():
    _check(lib.c2rt_cancel())


def render_device(scene_handle, cam, settings, d_rgb_ptr, d_argb_ptr=None, band=None, stream=None):
    """c2rt_render_device: device pointers (ints), async on `stream` (int cudaStream_t or None)."""
    b = C.byref(band) if band is not None else None
    _check(lib.c2rt_render_device(scene_handle, C.byref(cam), C.byref(settings), b, d_rgb_ptr, d_argb_ptr, stream, None))


def read_ray_counters(scene_handle, stream=None):
    p, s = C.c_uint64(), C.c_uint64()
    _check(lib.c2rt_read_ray_counters(scene_handle, stream, C.byref(p), C.byref(s)))
    return p.value, s.value


def deinterleave(gathered_ptr, frame_ptr, width, height, elem_words, n_ranks, band_rows, rows_pad, stream=None):
    _check(lib.c2rt_deinterleave(gathered_ptr, frame_ptr, width, height, elem_words, n_ranks, band_rows, rows_pad, stream))
