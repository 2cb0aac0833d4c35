"""ctypes loader of libsvr_b200.so (the C ABI declared in include/svr_abi.h).

Fails loudly if the library is missing or cannot be loaded: there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVR_B200_LIB", os.path.join(HERE, "csrc", "libsvr_b200.so"))
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "svr_abi.h")
HEADER_PATHS = [HEADER_PATH, os.path.join(os.path.dirname(HERE), "include", "pvr_abi.h")]

_lib = None

vp, ip, fp, dp = C.c_void_p, C.c_int, C.c_float, C.c_double
F = C.POINTER(C.c_float)
I = C.POINTER(C.c_int)
D = C.POINTER(C.c_double)
U8 = C.POINTER(C.c_ubyte)

_SIGNATURES = {
    "svr_create": (ip, [C.POINTER(vp), ip]),
    "svr_destroy": (ip, [vp]),
    "svr_last_error": (C.c_char_p, [vp]),
    "svr_abi_version": (ip, []),
    "svr_set_stream": (ip, [vp, vp]),
    "svr_synchronize": (ip, [vp]),
    "svr_launch_count": (C.c_int64, [vp]),
    "svr_set_tuning": (ip, [vp, ip, ip]),
    "svr_get_stream": (ip, [vp, C.POINTER(vp)]),
    "svr_set_async": (ip, [vp, ip]),
    "svr_make_current": (ip, [vp]),
    "svr_host_partition_strided": (ip, [ip, vp, ip, ip, vp, I]),
    "svr_init_reconstruction_volume": (ip, [vp, ip, ip, ip, fp, fp, fp, vp]),
    "svr_set_mask": (ip, [vp, ip, ip, ip, vp]),
    "svr_init_storage_volumes": (ip, [vp, ip, ip, ip]),
    "svr_fill_slices": (ip, [vp, vp, vp, vp]),
    "svr_set_slice_dims": (ip, [vp, vp, fp]),
    "svr_set_slice_matrices": (ip, [vp, vp, vp, vp, vp, vp, vp]),
    "svr_generate_psf_volume": (ip, [vp, vp, vp, fp]),
    "svr_update_scale_vector": (ip, [vp, vp, vp]),
    "svr_update_slice_weights": (ip, [vp, vp]),
    "svr_update_reconstructed": (ip, [vp, vp]),
    "svr_initialize_em_values": (ip, [vp]),
    "svr_gaussian_reconstruction": (ip, [vp, vp]),
    "svr_simulate_slices": (ip, [vp, vp]),
    "svr_initialize_robust_statistics": (ip, [vp, F]),
    "svr_estep": (ip, [vp, fp, fp, fp, vp]),
    "svr_mstep": (ip, [vp, ip, fp, F, F, F]),
    "svr_calculate_scale_vector": (ip, [vp, vp]),
    "svr_superresolution": (ip, [vp, ip, vp, ip, fp, fp, fp, fp, fp]),
    "svr_mask_volume": (ip, [vp]),
    "svr_scale_volume": (ip, [vp, F]),
    "svr_restore_slice_intensities": (ip, [vp, vp, ip, vp]),
    "svr_sync_cpu": (ip, [vp, vp]),
    "svr_get_vol_weights": (ip, [vp, vp]),
    "svr_debug_get": (ip, [vp, ip, vp]),
    "svr_gaussian_reconstruction_local": (ip, [vp]),
    "svr_gaussian_reconstruction_finish": (ip, [vp, vp]),
    "svr_superresolution_local": (ip, [vp, vp]),
    "svr_superresolution_finish": (ip, [vp, ip, fp, fp, fp, fp, fp]),
    "svr_mstep_local": (ip, [vp, D]),
    "svr_mstep_finish": (ip, [D, ip, fp, F, F, F]),
    "svr_initialize_robust_statistics_local": (ip, [vp, D]),
    "svr_scale_volume_local": (ip, [vp, D]),
    "svr_scale_volume_apply": (ip, [vp, fp]),
    "svr_device_buffer": (ip, [vp, ip, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "svr_profile_enable": (ip, [vp, ip]),
    "svr_profile_read": (ip, [vp, ip, D, C.POINTER(C.c_int64)]),
    "svr_profile_reset": (ip, [vp]),
    "svr_reg_init_storage": (ip, [vp, ip, ip, ip, fp, fp, fp]),
    "svr_reg_fill_slices": (ip, [vp, vp, vp]),
    "svr_reg_resample_slices": (ip, [vp, vp, vp, vp, vp]),
    "svr_reg_update_slices_i2w": (ip, [vp, vp]),
    "svr_reg_prepare": (ip, [vp]),
    "svr_reg_set_schedule": (ip, [vp, ip, ip, ip]),
    "svr_reg_register": (ip, [vp, vp]),
    "svr_reg_evaluate": (ip, [vp, vp, ip, vp]),
    "svr_reg_evaluations": (C.c_int64, [vp]),
    "svr_reg_debug_get": (ip, [vp, ip, vp]),
    "pvr_create": (ip, [C.POINTER(vp), ip]),
    "pvr_recon_init": (ip, [vp, ip, ip, ip, fp, fp, fp, vp, vp]),
    "pvr_recon_set_mask": (ip, [vp, vp]),
    "pvr_recon_reset": (ip, [vp]),
    "pvr_recon_reset_addon_cmap": (ip, [vp]),
    "pvr_recon_equalize": (ip, [vp]),
    "pvr_recon_copy_from_host": (ip, [vp, vp]),
    "pvr_recon_copy_to_host": (ip, [vp, vp]),
    "pvr_patches_init": (ip, [vp, ip, ip, ip, vp, vp]),
    "pvr_patches_set_matrices": (ip, [vp, vp, vp, vp, vp]),
    "pvr_patches_set_spx_masks": (ip, [vp, vp, ip]),
    "pvr_patches_copy_from_host": (ip, [vp, vp]),
    "pvr_patches_copy_to_host": (ip, [vp, vp]),
    "pvr_set_psf": (ip, [vp, vp, vp, fp]),
    "pvr_init_patch_based_recon": (ip, [vp, ip, vp, ip, ip, ip, vp]),
    "pvr_psf_reconstruction": (ip, [vp]),
    "pvr_psf_reconstruction_local": (ip, [vp]),
    "pvr_psf_reconstruction_finish": (ip, [vp]),
    "pvr_simulate_patches": (ip, [vp]),
    "pvr_superresolution_run": (ip, [vp]),
    "pvr_superresolution_regularize": (ip, [vp, ip, fp, fp, fp, fp, fp]),
    "pvr_rs_initialize_em_values": (ip, [vp]),
    "pvr_rs_initialize_robust_statistics": (ip, [vp, F]),
    "pvr_rs_estep_device": (ip, [vp, fp, fp, fp, vp]),
    "pvr_rs_get_scales_weights": (ip, [vp, vp, vp]),
    "pvr_rs_set_scales_weights": (ip, [vp, vp, vp]),
    "pvr_host_patch_em": (ip, [ip, vp, vp, vp, vp, fp, vp, vp]),
    "pvr_rs_mstep": (ip, [vp, ip, fp, F, F, F]),
    "pvr_rs_scale": (ip, [vp, vp]),
    "pvr_debug_get": (ip, [vp, ip, vp]),
    "svr_rreg_register": (ip, [vp, ip, ip, vp, vp, vp, vp, ip, vp, vp, vp, ip, C.c_size_t, vp, vp, vp, vp]),
    "svr_rreg_blur_with_padding": (ip, [vp, vp, vp, dp, ip, vp]),
    "svr_rreg_resample_with_padding": (ip, [vp, vp, vp, dp, dp, dp, ip, vp, C.c_size_t, vp]),
    "svr_host_slice_em": (ip, [ip, vp, vp, vp, vp, ip, vp, ip, dp, vp]),
    "svr_host_small_slices": (ip, [ip, vp, vp, I]),
    "svr_host_partition": (ip, [ip, vp, ip, ip, I, I]),
}


def declared_symbols() -> list[str]:
    """Every function include/svr_abi.h declares (used by the CPU test that checks the exports)."""
    text = ""
    for path in HEADER_PATHS:
        with open(path) as f:
            text += f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:svr|pvr)_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen the library and attach the prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m fetalreconstruction_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
