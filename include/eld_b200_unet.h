/* eld_b200_unet.h - U-Net part of the C ABI (included by eld_b200.h). */
#ifndef ELD_B200_UNET_H
#define ELD_B200_UNET_H
/* filled in below as the U-Net kernels land */
#endif
