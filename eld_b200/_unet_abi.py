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
    declare_engine(lib)


def declare_engine(lib):
    f32, sz = c.c_float, c.c_size_t
    lib.eld_conv3x3_wgrad_bf16.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.eld_deconv2x2_wgrad_bf16.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.eld_unet_param_count.restype = sz
    lib.eld_unet_param_offset.argtypes = [c.c_char_p, i32, c.POINTER(sz), c.POINTER(sz)]
    lib.eld_unet_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.eld_unet_workspace_bytes.restype = sz
    lib.eld_unet_create.argtypes = [vp, i32, i32, i32, i32, vp, sz, c.POINTER(vp)]
    lib.eld_unet_param_count_io.argtypes = [i32, i32]
    lib.eld_unet_param_count_io.restype = sz
    lib.eld_unet_param_offset_io.argtypes = [c.c_char_p, i32, i32, i32, c.POINTER(sz), c.POINTER(sz)]
    lib.eld_unet_create_io.argtypes = [vp, i32, i32, i32, i32, vp, sz, i32, i32, c.POINTER(vp)]
    lib.eld_unet_grad_buckets_io.argtypes = [i32, i32, c.POINTER(sz), i32]
    lib.eld_unet_destroy.argtypes = [vp]
    lib.eld_unet_destroy.restype = None
    lib.eld_unet_forward.argtypes = [vp, vp, vp, vp, vp]
    lib.eld_unet_train_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.eld_adam_step.argtypes = [vp, vp, vp, vp, vp, sz, f32, f32, f32, f32, f32, i32, f32, vp]
    lib.eld_unet_grad_buckets.argtypes = [c.POINTER(sz), i32]
    lib.eld_unet_bucket_events.argtypes = [vp, i32]
    lib.eld_unet_wait_bucket.argtypes = [vp, i32, vp]
    lib.eld_unet_backward.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.eld_unet_set_loss.argtypes = [vp, i32]
    lib.eld_clock_probe.argtypes = [vp, vp, vp]
    lib.eld_unet_profile.argtypes = [vp, i32]
    lib.eld_unet_profile_read.argtypes = [vp, i32, c.c_char_p, vp, vp, vp, c.POINTER(i32)]
