"""ctypes driver for the CPU oracle (oracle/meao_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/meao_oracle.h.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.  PARITY UNPINNED:
the reference has no golden vectors for this path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OracleParams(C.Structure):
    _fields_ = [("noise_filter_tolerance", C.c_float), ("blur_tolerance", C.c_float),
                ("upsample_tolerance", C.c_float), ("thickness_modifier", C.c_float),
                ("intensity", C.c_float)]


class OracleCamera(C.Structure):
    _fields_ = [("near_clip", C.c_float), ("far_clip", C.c_float),
                ("tan_half_fov_h", C.c_float), ("reversed_z", C.c_int)]


class _OracleStruct(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("lw", C.c_int * 7), ("lh", C.c_int * 7),
                ("quantize_storage", C.c_int), ("depth_is_linear", C.c_int),
                ("params", OracleParams), ("camera", OracleCamera),
                ("sample_exhaustively", C.c_int), ("single_pass_stereo", C.c_int), ("high_quality_mask", C.c_int),
                ("single_scale", C.c_int)]
    # buffer pointers follow; they are reached through meao_oracle_get_buffer


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc; a few seconds)."""
    so = os.path.join(_HERE, "libmeao_oracle.so")
    src = os.path.join(_HERE, "meao_oracle.c")
    hdr = os.path.join(_HERE, "meao_oracle.h")
    mk = os.path.join(_HERE, "Makefile")
    stale = (not os.path.exists(so)
             or any(not os.path.exists(os.path.join(_HERE, f"libmeao_oracle_{v}.so")) for v in VARIANTS if v != "fma")
             or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(mk)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)


# "fma" is THE oracle.  The others restate conventions the reference leaves open (meao_oracle.h) and are used only to measure
# how often each alternative would flip an output code: unfused mad, truncating f16 stores, x/y as x*(1/y), truncated quotients.
VARIANTS = ("fma", "nofma", "f16rtz", "divmulrcp", "divrtz")
_libs: dict[str, C.CDLL] = {}


def _lib(variant: str = "fma") -> C.CDLL:
    if variant not in _libs:
        build()
        assert variant in VARIANTS, variant
        name = "libmeao_oracle.so" if variant == "fma" else f"libmeao_oracle_{variant}.so"
        lib = C.CDLL(os.path.join(_HERE, name))
        lib.meao_oracle_create.restype = C.POINTER(_OracleStruct)
        lib.meao_oracle_create.argtypes = [C.c_int, C.c_int]
        lib.meao_oracle_destroy.argtypes = [C.POINTER(_OracleStruct)]
        lib.meao_oracle_get_buffer.restype = C.POINTER(C.c_float)
        lib.meao_oracle_get_buffer.argtypes = [C.POINTER(_OracleStruct), C.c_int,
                                               C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.meao_oracle_buffer_mut.restype = C.POINTER(C.c_float)
        lib.meao_oracle_buffer_mut.argtypes = lib.meao_oracle_get_buffer.argtypes
        fp = C.POINTER(C.c_float)
        lib.meao_oracle_downsample.argtypes = [C.POINTER(_OracleStruct), fp, C.c_int]
        lib.meao_oracle_render.argtypes = [C.POINTER(_OracleStruct), C.c_int, C.c_int]
        lib.meao_oracle_upsample.argtypes = [C.POINTER(_OracleStruct), C.c_int, C.c_int]
        lib.meao_oracle_render_wide.argtypes = [C.POINTER(_OracleStruct), C.c_int, C.c_int]
        lib.meao_oracle_debug_view.argtypes = [C.POINTER(_OracleStruct), C.c_int, C.c_void_p]
        lib.meao_oracle_run.argtypes = [C.POINTER(_OracleStruct), fp, C.c_int]
        lib.meao_oracle_f16_round.restype = C.c_float
        lib.meao_oracle_f16_round.argtypes = [C.c_float]
        lib.meao_oracle_f32_to_f16_bits.restype = C.c_uint16
        lib.meao_oracle_f32_to_f16_bits.argtypes = [C.c_float]
        lib.meao_oracle_f16_bits_to_f32.restype = C.c_float
        lib.meao_oracle_f16_bits_to_f32.argtypes = [C.c_uint16]
        lib.meao_oracle_unorm8_code.restype = C.c_uint8
        lib.meao_oracle_unorm8_code.argtypes = [C.c_float]
        lib.meao_oracle_zbuffer_params.argtypes = [C.POINTER(OracleCamera), fp]
        lib.meao_oracle_sample_thickness.argtypes = [fp]
        lib.meao_oracle_render_constants.argtypes = [C.POINTER(_OracleStruct), C.c_int, fp, fp, fp, fp, fp]
        lib.meao_oracle_render_constants_wide.argtypes = [C.POINTER(_OracleStruct), C.c_int, fp, fp, fp, fp, fp]
        lib.meao_oracle_upsample_constants.argtypes = [C.POINTER(_OracleStruct), C.c_int, fp, fp, fp, fp, fp, fp]
        lib.meao_oracle_composite_framebuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
        lib.meao_oracle_composite_gbuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
        lib.meao_oracle_composite_debug.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
        _libs[variant] = lib
    return _libs[variant]


def tan_half_fov_h(width: int, height: int, fov_y_deg: float = 60.0) -> float:
    """1 / projectionMatrix[0,0] = aspect * tan(fovY / 2)  (AmbientOcclusion.cs:570-573)."""
    return float(np.float32(width / height * np.tan(np.radians(fov_y_deg) / 2.0)))


class Oracle:
    """One oracle instance per resolution; mirrors the AmbientOcclusion parameter surface."""

    BUFFER_NAMES = {1: "LinearDepth", 2: "LowDepth1", 3: "LowDepth2", 4: "LowDepth3", 5: "LowDepth4",
                    6: "TiledDepth1", 7: "TiledDepth2", 8: "TiledDepth3", 9: "TiledDepth4",
                    10: "Occlusion1", 11: "Occlusion2", 12: "Occlusion3", 13: "Occlusion4",
                    14: "Combined1", 15: "Combined2", 16: "Combined3", 17: "AmbientOcclusion",
                    # extension ids (not in AO.cs:787-808): output of Render.compute kernel "main" per level
                    18: "HighQuality1", 19: "HighQuality2", 20: "HighQuality3", 21: "HighQuality4"}

    def __init__(self, width: int, height: int, *, variant: str = "fma", quantize_storage: bool = True,
                 depth_is_linear: bool = False, near: float = 0.3, far: float = 100.0,
                 tan_half_fov_h_: float | None = None, reversed_z: bool = True, threads: int = 1,
                 noise_filter_tolerance: float = 0.0, blur_tolerance: float = -4.6,
                 upsample_tolerance: float = -12.0, thickness_modifier: float = 1.0, intensity: float = 1.0,
                 sample_exhaustively: bool = False, single_pass_stereo: bool = False, high_quality_mask: int = 0,
                 single_scale: bool = False):
        self._lib = _lib(variant)
        self._o = self._lib.meao_oracle_create(width, height)
        if not self._o:
            raise ValueError("bad dimensions")
        self.width, self.height, self.threads = width, height, threads
        s = self._o.contents
        s.quantize_storage = int(quantize_storage)
        s.depth_is_linear = int(depth_is_linear)
        s.camera.near_clip, s.camera.far_clip = near, far
        s.camera.tan_half_fov_h = tan_half_fov_h(width, height) if tan_half_fov_h_ is None else tan_half_fov_h_
        s.camera.reversed_z = int(reversed_z)
        s.sample_exhaustively, s.single_pass_stereo, s.high_quality_mask = int(sample_exhaustively), int(single_pass_stereo), int(high_quality_mask)
        s.single_scale = int(single_scale)
        p = s.params
        p.noise_filter_tolerance, p.blur_tolerance, p.upsample_tolerance = noise_filter_tolerance, blur_tolerance, upsample_tolerance
        p.thickness_modifier, p.intensity = thickness_modifier, intensity

    def __del__(self):
        if getattr(self, "_o", None):
            self._lib.meao_oracle_destroy(self._o)
            self._o = None

    @property
    def params(self) -> OracleParams:
        return self._o.contents.params

    @property
    def camera(self) -> OracleCamera:
        return self._o.contents.camera

    def level_dims(self, level: int) -> tuple[int, int]:
        s = self._o.contents
        return s.lw[level], s.lh[level]

    @staticmethod
    def _fptr(a: np.ndarray):
        return a.ctypes.data_as(C.POINTER(C.c_float))

    def _depth(self, depth: np.ndarray) -> np.ndarray:
        d = np.ascontiguousarray(depth, dtype=np.float32)
        assert d.shape == (self.height, self.width), d.shape
        return d

    # ---- stages -----------------------------------------------------------------------------
    def downsample(self, depth: np.ndarray) -> None:
        d = self._depth(depth)
        self._lib.meao_oracle_downsample(self._o, self._fptr(d), self.threads)

    def render(self, level: int) -> None:
        self._lib.meao_oracle_render(self._o, level, self.threads)

    def render_wide(self, level: int) -> None:
        """Render.compute kernel "main" (WIDE_SAMPLING) on LowDepth<level> -> HighQuality<level> (buffer 17 + level)."""
        self._lib.meao_oracle_render_wide(self._o, level, self.threads)

    def upsample(self, lo_level: int) -> None:
        self._lib.meao_oracle_upsample(self._o, lo_level, self.threads)

    def debug_view(self, debug_id: int) -> np.ndarray:
        """The W x H R8 image PushDebugBlitCommands (AO.cs:787-820) leaves in _result for buffer <debug_id>."""
        out = np.zeros((self.height, self.width), np.uint8)
        self._lib.meao_oracle_debug_view(self._o, debug_id, out.ctypes.data)
        return out

    def run(self, depth: np.ndarray) -> np.ndarray:
        d = self._depth(depth)
        self._lib.meao_oracle_run(self._o, self._fptr(d), self.threads)
        return self.ao_u8()

    # ---- buffers ----------------------------------------------------------------------------
    def buffer(self, debug_id: int) -> np.ndarray:
        """Float view (post-quantisation values) of debug buffer 1..17; tiled -> [16, h, w]."""
        w, h, s = C.c_int(), C.c_int(), C.c_int()
        p = self._lib.meao_oracle_get_buffer(self._o, debug_id, C.byref(w), C.byref(h), C.byref(s))
        if not p:
            raise KeyError(debug_id)
        n = w.value * h.value * s.value
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a.reshape((s.value, h.value, w.value)) if s.value > 1 else a.reshape((h.value, w.value))

    def set_buffer(self, debug_id: int, values: np.ndarray) -> None:
        self.buffer(debug_id)[...] = np.asarray(values, dtype=np.float32)

    def codes(self, debug_id: int) -> np.ndarray:
        """UNORM8 buffer (ids 10..21) as uint8 codes."""
        assert 10 <= debug_id <= 21
        return np.rint(self.buffer(debug_id) * 255.0).astype(np.uint8)

    def ao_u8(self) -> np.ndarray:
        return self.codes(17)

    # ---- constants --------------------------------------------------------------------------
    def zbuffer_params(self) -> np.ndarray:
        out = np.zeros(4, np.float32)
        self._lib.meao_oracle_zbuffer_params(C.byref(self._o.contents.camera), self._fptr(out))
        return out

    def sample_thickness(self) -> np.ndarray:
        out = np.zeros(12, np.float32)
        self._lib.meao_oracle_sample_thickness(self._fptr(out))
        return out

    def render_constants(self, level: int, wide: bool = False) -> dict:
        it, sw, isd = np.zeros(12, np.float32), np.zeros(12, np.float32), np.zeros(2, np.float32)
        rf, inten = C.c_float(), C.c_float()
        fn = self._lib.meao_oracle_render_constants_wide if wide else self._lib.meao_oracle_render_constants
        fn(self._o, level, self._fptr(it), self._fptr(sw), self._fptr(isd),
                                               C.byref(rf), C.byref(inten))
        return {"inv_thickness": it, "sample_weight": sw, "inv_slice_dim": isd,
                "reject_fadeoff": np.float32(rf.value), "intensity": np.float32(inten.value)}

    def upsample_constants(self, lo_level: int) -> dict:
        il, ih = np.zeros(2, np.float32), np.zeros(2, np.float32)
        nfs, ss, bt, ut = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        self._lib.meao_oracle_upsample_constants(self._o, lo_level, self._fptr(il), self._fptr(ih),
                                                 C.byref(nfs), C.byref(ss), C.byref(bt), C.byref(ut))
        return {"inv_low": il, "inv_high": ih, "noise_filter_strength": np.float32(nfs.value),
                "step_size": np.float32(ss.value), "blur_tolerance": np.float32(bt.value),
                "upsample_tolerance": np.float32(ut.value)}

    # ---- scalar conversions -------------------------------------------------------------------
    def f16_round(self, x: float) -> float:
        return self._lib.meao_oracle_f16_round(C.c_float(x))

    def f16_bits(self, x: float) -> int:
        return self._lib.meao_oracle_f32_to_f16_bits(C.c_float(x))

    def unorm8_code(self, x: float) -> int:
        return self._lib.meao_oracle_unorm8_code(C.c_float(x))


def composite_framebuffer(ao_codes: np.ndarray, color: np.ndarray) -> np.ndarray:
    """Blit.shader pass 2: color (uint8 [...,4] RGBA8 or float16 [...,4] RGBA16F) *= ao.  Returns a new array."""
    ao = np.ascontiguousarray(ao_codes, np.uint8)
    out = np.ascontiguousarray(color).copy()
    assert out.dtype in (np.uint8, np.float16) and out.shape[-1] == 4 and out.size == ao.size * 4
    _lib().meao_oracle_composite_framebuffer(ao.ctypes.data, out.ctypes.data, int(out.dtype == np.float16), ao.size)
    return out


def composite_gbuffer(ao_codes: np.ndarray, gbuffer0: np.ndarray, gbuffer3: np.ndarray):
    """Blit.shader pass 1: gbuffer0 (RGBA8).a and gbuffer3 (RGBA8 / RGBA16F).rgb *= 1-(1-ao).  Returns new arrays."""
    ao = np.ascontiguousarray(ao_codes, np.uint8)
    g0 = np.ascontiguousarray(gbuffer0).copy()
    g3 = np.ascontiguousarray(gbuffer3).copy()
    assert g0.dtype == np.uint8 and g3.dtype in (np.uint8, np.float16)
    _lib().meao_oracle_composite_gbuffer(ao.ctypes.data, g0.ctypes.data, g3.ctypes.data, int(g3.dtype == np.float16), ao.size)
    return g0, g3


def composite_debug(view_codes: np.ndarray, like: np.ndarray) -> np.ndarray:
    """Blit.shader pass 3: a new [..., 4] array of like's dtype (uint8 RGBA8 / float16 RGBA16F) = view.rrrr."""
    v = np.ascontiguousarray(view_codes, np.uint8)
    out = np.zeros(v.shape + (4,), like.dtype)
    assert out.dtype in (np.uint8, np.float16)
    _lib().meao_oracle_composite_debug(v.ctypes.data, out.ctypes.data, int(out.dtype == np.float16), v.size)
    return out
