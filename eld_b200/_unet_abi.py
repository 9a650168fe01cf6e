"""ctypes prototypes of the U-Net entry points (include/eld_b200_unet.h)."""


def declare(lib):
    pass
