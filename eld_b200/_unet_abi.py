"""ctypes prototypes of the U-Net entry points (include/eld_b200_unet.h)."""
import ctypes as c

vp, i32 = c.c_void_p, c.c_int


def declare(lib):
    lib.eld_pack_weights.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.eld_conv3x3_bf16.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                     i32, vp, i32, i32, vp]
    lib.eld_deconv2x2_bf16.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.eld_deconv2x2_dgrad_bf16.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32,
                                             i32, vp, i32, i32, vp]
