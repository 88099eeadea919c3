"""ctypes binding of the C-ABI in include/rgbd_b200.h.

The product path has no CPU fallback: if librgbd_b200.so is missing this module raises on
first use (build it with `python -m __graft_entry__` / build.py).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librgbd_b200.so")

DT_F32, DT_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_GELU = 0, 1, 2, 3
EPI_LINEAR, EPI_GATE, EPI_BILERP, EPI_SHUFFLE2 = 0, 1, 2, 3
MAX_TAPS = 25


class RgbdError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("y2", C.c_void_p), ("w", C.c_void_p),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("mul", C.c_void_p), ("in_scale", C.c_void_p),
        ("sched_ws", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("x_cstride", C.c_int32), ("x_coff", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("Cout", C.c_int32), ("y_cstride", C.c_int32), ("y_coff", C.c_int32),
        ("y2_cstride", C.c_int32), ("y2_coff", C.c_int32),
        ("Hs", C.c_int32), ("Ws", C.c_int32), ("o_step", C.c_int32), ("o_off_y", C.c_int32),
        ("o_off_x", C.c_int32), ("i_step", C.c_int32),
        ("ntaps", C.c_int32),
        ("res_cstride", C.c_int32), ("res_coff", C.c_int32), ("res_H", C.c_int32), ("res_W", C.c_int32),
        ("mul_cstride", C.c_int32), ("mul_coff", C.c_int32),
        ("act", C.c_int32), ("epi", C.c_int32),
        ("x_dtype", C.c_int32), ("y_dtype", C.c_int32),
        ("cout_pad", C.c_int32), ("w_image_stride", C.c_int32),
        ("dy", C.c_int8 * MAX_TAPS), ("dx", C.c_int8 * MAX_TAPS), ("wtap", C.c_int8 * MAX_TAPS),
        ("_pad", C.c_int8 * 1),
    ]


class RbDesc(C.Structure):
    """rgbd_rb_desc (include/rgbd_b200.h)"""
    _fields_ = [
        ("x", C.c_void_p), ("res", C.c_void_p), ("y", C.c_void_p),
        ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p),
        ("b1", C.c_void_p), ("b2", C.c_void_p), ("b3", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("x_cstride", C.c_int32), ("x_coff", C.c_int32),
        ("Cmid", C.c_int32), ("Cout", C.c_int32),
        ("res_cstride", C.c_int32), ("res_coff", C.c_int32), ("y_cstride", C.c_int32), ("y_coff", C.c_int32),
        ("final_relu", C.c_int32), ("_pad", C.c_int32), ("sched_ws", C.c_void_p),
    ]


class RansTables(C.Structure):
    _fields_ = [
        ("cdf", C.c_void_p), ("base", C.c_void_p), ("length", C.c_void_p), ("offset", C.c_void_p),
        ("n_tables", C.c_int32), ("total", C.c_int32), ("enc_rec", C.c_void_p),
    ]


_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes; every function returns int except where noted
_PROTOS = {
    "rgbd_conv_simt": [C.POINTER(ConvDesc), _vp],
    "rgbd_conv_validate": [C.POINTER(ConvDesc)],
    "rgbd_conv_tc_plan_create": [C.POINTER(ConvDesc), _i32, C.POINTER(_vp)],
    "rgbd_conv_tc_run": [_vp, _vp],
    "rgbd_rb_plan_create": [C.POINTER(RbDesc), C.POINTER(_vp)],
    "rgbd_rb_run": [_vp, _vp],
    "rgbd_se_scale": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _vp, _vp],
    "rgbd_se_partial": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "rgbd_se_gate": [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp],
    "rgbd_maxpool7s3": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_nchw_to_nhwc": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_nhwc_to_nchw": [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_scale_channels": [_vp, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_scale_weights": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_copy_view": [_vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_cast_view_bf16": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_zero": [_vp, _i64, _vp],
    "rgbd_ckbd_quantize_index": [_vp, _i32, _i32, _vp, _vp, _i32, _f32, _i32, _i32, _i32, _i32, _i32, _vp, _vp,
                                 _i64, _i64, _vp, _i32, _i32, _i32, _vp],
    "rgbd_ckbd_index": [_vp, _vp, _i32, _f32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _vp],
    "rgbd_ckbd_dequant_scatter": [_vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp],
    "rgbd_ckbd_ste_likelihood": [_vp, _i32, _i32, _vp, _f32, _f32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32,
                                 _i32, _vp, _i32, _i32, _vp],
    "rgbd_eb_quantize": [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "rgbd_eb_dequantize": [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp],
    "rgbd_eb_likelihood": [_vp, _i32, _i32, _i32, _i32, _vp, _f32, _vp, _i32, _i32, _i32, _vp, _vp],
    "rgbd_rans_encode": [_vp, _vp, _i64, _i32, _i32, C.POINTER(RansTables), _vp, _i64, _vp, _vp],
    "rgbd_gather_streams": [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _vp],
    "rgbd_rans_decode_init": [_vp, _vp, _i32, _vp, _vp],
    "rgbd_rans_decode_chunk": [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i32, C.POINTER(RansTables), _vp],
    "rgbd_rans_decode_streams": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i32,
                                 C.POINTER(RansTables), _vp],
    "rgbd_layernorm": [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _f32, _i32, _i32, _i32, _vp],
    "rgbd_pixel_shuffle2": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_window_attention": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _f32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_sq_error_sums": [_vp, _vp, _i32, _i64, _i32, _vp, _i32, _vp, _vp],
    "rgbd_ssim_level": [_vp, _vp, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp],
    "rgbd_avgpool2": [_vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "rgbd_quantize_u8": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "rgbd_quantize_u16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "rgbd_pmf_to_quantized_cdf": [_vp, _i32, _i32, _vp],
}

# every symbol include/rgbd_b200.h declares (checked by tests/test_abi.py)
EXPORTS = sorted(list(_PROTOS.keys()) + ["rgbd_last_error", "rgbd_abi_version", "rgbd_launch_count", "rgbd_count_launch",
                                          "rgbd_conv_tc_plan_destroy", "rgbd_rb_plan_destroy", "rgbd_ssim_work_elems"])
EXPORTS.remove("rgbd_conv_validate")

_lib = None


def load():
    """Load librgbd_b200.so (once). Raises RgbdError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RgbdError(
            f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.rgbd_last_error.restype = C.c_char_p
    lib.rgbd_last_error.argtypes = []
    lib.rgbd_abi_version.restype = C.c_int
    lib.rgbd_launch_count.restype = C.c_int64
    lib.rgbd_launch_count.argtypes = [C.c_int]
    lib.rgbd_count_launch.restype = None
    lib.rgbd_count_launch.argtypes = [C.c_int]
    lib.rgbd_conv_tc_plan_destroy.restype = None
    lib.rgbd_conv_tc_plan_destroy.argtypes = [_vp]
    lib.rgbd_ssim_work_elems.restype = C.c_int64
    lib.rgbd_ssim_work_elems.argtypes = [_i32, _i32, _i32]
    lib.rgbd_rb_plan_destroy.restype = None
    lib.rgbd_rb_plan_destroy.argtypes = [_vp]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().rgbd_last_error().decode("utf-8", "replace")
        raise RgbdError(f"{what} failed with code {rc}: {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
