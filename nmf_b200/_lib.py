"""ctypes binding of libnmf_b200.so (the C ABI declared in include/nmf_b200.h).

The product path has no CPU fallback: if the CUDA library is missing, `lib()` raises.  Build it with
`python -m nmf_b200.build` (or `__graft_entry__.build()`); the .so lives in-tree next to this file.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnmf_b200.so")

c_float3 = C.c_float * 3
c_int3 = C.c_int * 3
c_ptr3 = C.c_void_p * 3


class NmfScene(C.Structure):
    _fields_ = [
        ("aabb0", c_float3), ("aabb1", c_float3), ("inv_aabb2", c_float3),
        ("stepsize", C.c_float), ("near", C.c_float), ("far", C.c_float),
        ("distance_scale", C.c_float), ("density_shift", C.c_float),
        ("n_steps", C.c_int),
        ("plane_w", c_int3), ("plane_h", c_int3), ("line_n", c_int3),
        ("occ_vox", C.c_void_p), ("occ_cell", C.c_void_p),
        ("ow", C.c_int), ("oh", C.c_int), ("od", C.c_int), ("opitch", C.c_int), ("has_occ", C.c_int),
        ("occ_coarse", C.c_void_p), ("ocw", C.c_int), ("och", C.c_int), ("ocd", C.c_int), ("occ_scale", c_float3),
        ("dval", c_ptr3), ("dpack", c_ptr3), ("lval", c_ptr3), ("lpack", c_ptr3),
        ("aval", c_ptr3), ("alval", c_ptr3), ("basis_t", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p),
        ("diffuse_mul", C.c_float), ("diffuse_bias", C.c_float), ("tint_bias", C.c_float),
        ("f0_bias", C.c_float), ("roughness_bias", C.c_float),
        ("brdf_w0t", C.c_void_p), ("brdf_b0", C.c_void_p), ("brdf_w1t", C.c_void_p), ("brdf_b1", C.c_void_p),
        ("brdf_w2t", C.c_void_p), ("brdf_b2", C.c_void_p),
        ("brdf_bias", C.c_float), ("anoise", C.c_float),
        ("sobol", C.c_void_p), ("sh_conv", C.c_void_p),
        ("env_sat", C.c_void_p), ("env_h", C.c_int), ("env_w", C.c_int), ("env_mipbias", C.c_float),
        ("env_top", c_float3), ("env_bot", c_float3),
        ("plain_w0t", C.c_void_p), ("plain_b0", C.c_void_p), ("plain_w1t", C.c_void_p), ("plain_b1", C.c_void_p),
        ("plain_w2t", C.c_void_p), ("plain_b2", C.c_void_p),
        ("plain_w0", C.c_void_p), ("plain_w1", C.c_void_p),
        ("rays_per_ray", C.c_int), ("max_brdf_rays1", C.c_int), ("max_retrace", C.c_int), ("model", C.c_int),
        ("brdf_w0u", C.c_void_p), ("brdf_w1u", C.c_void_p), ("brdf_w2u", C.c_void_p), ("mlp_mode", C.c_int),
        ("brdf_w0b", C.c_void_p), ("brdf_w1b", C.c_void_p), ("brdf_w2b", C.c_void_p),
        ("env_sat2", C.c_void_p), ("env_dyn", C.c_void_p),
    ]


class NmfRender(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int), ("chunk", C.c_int), ("focal", C.c_float),
        ("seed", C.c_uint64), ("ray_id0", C.c_uint64),
        ("skip_eps", C.c_float), ("t_cut", C.c_float), ("white_bg", C.c_int), ("cap_scale", C.c_float),
    ]


class NmfPlainGrads(C.Structure):
    _fields_ = [("d_plane", c_ptr3), ("d_line", c_ptr3), ("a_plane", c_ptr3), ("a_line", c_ptr3), ("basis_t", C.c_void_p),
                ("w0t", C.c_void_p), ("b0", C.c_void_p), ("w1t", C.c_void_p), ("b1", C.c_void_p),
                ("w2t", C.c_void_p), ("b2", C.c_void_p)]


class NmfNormalGrads(C.Structure):
    _fields_ = [("gpack", c_ptr3), ("glpack", c_ptr3)]


class NmfMicrofacetGrads(C.Structure):
    _fields_ = [("d_plane", c_ptr3), ("d_line", c_ptr3), ("a_plane", c_ptr3), ("a_line", c_ptr3), ("basis_t", C.c_void_p),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p),
                ("w0t", C.c_void_p), ("b0", C.c_void_p), ("w1t", C.c_void_p), ("b1", C.c_void_p),
                ("w2t", C.c_void_p), ("b2", C.c_void_p), ("gsat", C.c_void_p), ("d_mipbias", C.c_void_p),
                ("normals", NmfNormalGrads)]


class NmfShadingPack(C.Structure):
    _fields_ = [("w", C.c_void_p * 3), ("b", C.c_void_p * 3), ("n_out", C.c_int32 * 3), ("n_in", C.c_int32 * 3),
                ("head_w", C.c_void_p * 4), ("head_b", C.c_void_p * 4), ("head_rows", C.c_int32 * 4),
                ("wt", C.c_void_p * 3), ("bo", C.c_void_p * 3), ("w16", C.c_void_p * 3), ("wbf", C.c_void_p * 3),
                ("head_w_out", C.c_void_p), ("head_b_out", C.c_void_p)]


class NmfTransposeJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_uint64), ("c", C.c_int32), ("pad", C.c_int32)]


class NmfMicrofacetTrain(C.Structure):
    _fields_ = [("lambda_pred", C.c_float), ("lambda_ori", C.c_float), ("detach_N", C.c_int), ("loss", C.c_void_p)]


class NmfTrain(C.Structure):
    _fields_ = [("n_rays", C.c_int), ("focal", C.c_float), ("seed", C.c_uint64), ("ray_id0", C.c_uint64),
                ("ray_ids", C.c_void_p), ("max_samples", C.c_int), ("cap_samples", C.c_int),
                ("lambda_pred", C.c_float), ("white_bg", C.c_int)]


class NmfTrainOut(C.Structure):
    _fields_ = [("rgb_map", C.c_void_p), ("acc_map", C.c_void_p), ("whole_valid", C.c_void_p), ("loss", C.c_void_p),
                ("n_kept", C.c_void_p), ("error", C.c_void_p)]


class NmfAdam(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step", C.c_int), ("grad_scale", C.c_float), ("max_norm", C.c_float),
                ("control", C.c_void_p)]


class NmfRenderTrain(C.Structure):
    _fields_ = [("max_samples", C.c_int), ("min_rough", C.c_float), ("whole_valid", C.c_void_p), ("n_kept", C.c_void_p)]


IMAGE_FIELDS = ["rgb_map", "acc_map", "depth", "world_normal", "normal", "termination_xyz", "surf_width",
                "cross_section", "diffuse", "tint", "roughness", "spec", "albedo"]
COUNTER_FIELDS = ["n_samples0", "n_samples1", "n_cand", "n_bounce_rays0", "n_bounce_rays1", "n_retrace",
                  "n_shaded", "error", "stat4"]


class NmfImages(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in IMAGE_FIELDS]


class NmfCounters(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in COUNTER_FIELDS]


class NmfError(RuntimeError):
    pass


class NmfOverflow(NmfError):
    """A device-side scratch list was too small for the scene (NmfCounters.error): the render is incomplete."""


_ERRORS = {-1: "NMF_E_ARG (null pointer or non-positive size)",
           -2: "NMF_E_UNSUPPORTED (shape outside what the kernels are compiled for)",
           -3: "NMF_E_WORKSPACE (workspace too small)"}
N_PHASES = 12
DEV_ERRORS = {1: "surviving-sample list overflowed", 2: "bounce-sample list overflowed",
              4: "a chunk's bounce-ray region overflowed", 8: "the valid-sample list of the reverse pass overflowed"}


def check(status, what):
    if status == 0:
        return
    if status < 0:
        raise NmfError(f"{what}: {_ERRORS.get(status, status)}")
    raise NmfError(f"{what}: CUDA error {status}")


_lib = None


def lib():
    """Loads libnmf_b200.so; raises (loudly) when it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NmfError(f"{LIB_PATH} is missing: build it with `python -m nmf_b200.build` "
                       "(nmf_b200 has no CPU / PyTorch fallback for the render path)")
    L = C.CDLL(LIB_PATH)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    SP, RP = C.POINTER(NmfScene), C.POINTER(NmfRender)
    IP, CP = C.POINTER(NmfImages), C.POINTER(NmfCounters)
    sigs = {
        "nmf_abi_version": (I, []),
        "nmf_profile_enable": (I, [I]),
        "nmf_profile_read": (I, [P, I]),
        "nmf_profile_phase_name": (C.c_char_p, [I]),
        "nmf_workspace_bytes": (C.c_size_t, [SP, I, I]),
        "nmf_workspace_bytes_scaled": (C.c_size_t, [SP, I, I, F]),
        "nmf_render_rays": (I, [SP, RP, P, IP, CP, P, C.c_size_t, P]),
        "nmf_render_rays_host": (I, [SP, RP, P, P, IP, IP, CP, CP, P, C.c_size_t, P]),
        "nmf_sample_rays": (I, [SP, P, I, F, P, P, P, P]),
        "nmf_vm_density": (I, [SP, P, I, I, I, P, P]),
        "nmf_vm_appfeature": (I, [SP, P, I, I, P, P]),
        "nmf_vm_normals": (I, [SP, P, I, I, P, P]),
        "nmf_env_lookup": (I, [SP, P, P, I, P, P]),
        "nmf_ggx_sample": (I, [P, P, P, P, I, P, P, P, P, P]),
        "nmf_brdf_mlp": (I, [SP, P, P, P, P, I, P, P]),
        "nmf_material_heads": (I, [SP, P, I, P, P, P, P, P, P]),
        "nmf_dense_alpha": (I, [SP, I, I, I, P, P, P]),
        "nmf_generate_rays": (I, [P, I, I, F, F, F, F, P, I, P, P]),
        "nmf_image_sq_error": (I, [P, P, P, I, P, P]),
        "nmf_sample_rays_train": (I, [SP, P, I, F, C.c_uint64, C.c_uint64, P, I, P, P, P, P, P, P]),
        "nmf_train_workspace_bytes": (C.c_size_t, [SP, I, I]),
        "nmf_upsample_bilinear": (I, [P, I, I, I, P, I, I, P]),
        "nmf_l1_reg": (I, [P, C.c_size_t, F, P, P, P]),
        "nmf_grad_sq_norm": (I, [P, C.c_size_t, P, P]),
        "nmf_adam_step": (I, [P, P, P, P, C.c_size_t, C.POINTER(NmfAdam), P, P]),
        "nmf_env_lookup_bwd_scatter": (I, [SP, P, P, P, I, P, P]),
        "nmf_env_lookup_bwd_finish": (I, [P, I, I, P, F, F, P, P, P, P]),
        "nmf_env_lookup_bwd_mipbias": (I, [SP, P, P, P, I, P, P]),
        "nmf_vm_normals_bwd_scatter": (I, [SP, P, I, I, P, C.POINTER(NmfNormalGrads), P]),
        "nmf_vm_normals_bwd_finish": (I, [SP, C.POINTER(NmfNormalGrads), P, P, P, P, P]),
        "nmf_material_heads_bwd": (I, [SP, P, P, P, P, I, P, P, P, P]),
        "nmf_render_train_workspace_bytes": (C.c_size_t, [SP, I, F]),
        "nmf_render_rays_train": (I, [SP, RP, C.POINTER(NmfRenderTrain), P, IP, CP, P, C.c_size_t, P]),
        "nmf_train_microfacet": (I, [SP, RP, C.POINTER(NmfRenderTrain), C.POINTER(NmfMicrofacetTrain), P, P,
                                     C.POINTER(NmfMicrofacetGrads), IP, CP, P, C.c_size_t, P]),
        "nmf_pack_factor": (I, [P, I, I, I, P, P, P, P, P]),
        "nmf_env_build_sat": (I, [P, I, I, F, F, P, P, P, P, P]),
        "nmf_occupancy_from_alpha": (I, [P, I, I, I, F, I, P, P, P, P, P]),
        "nmf_bench_gather": (I, [P, C.c_size_t, I, I, I, P, P]),
        "nmf_transpose_batch": (I, [P, I, I, P]),
        "nmf_env_pair_sat": (I, [P, I, I, P, P]),
        "nmf_pack_shading": (I, [C.POINTER(NmfShadingPack), P]),
        "nmf_env_build_sat_dev": (I, [P, I, I, P, P, P, P, P, P, P]),
        "nmf_env_lookup_bwd_finish_dev": (I, [P, I, I, P, P, P, P, P, P]),
        "nmf_train_plain": (I, [SP, C.POINTER(NmfTrain), P, P, C.POINTER(NmfPlainGrads), C.POINTER(NmfTrainOut), P,
                                C.c_size_t, P]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    assert L.nmf_abi_version() == 10, "libnmf_b200.so ABI mismatch: rebuild"
    _lib = L
    return L


EXPORTED = ["nmf_abi_version", "nmf_profile_enable", "nmf_profile_read", "nmf_profile_phase_name", "nmf_workspace_bytes", "nmf_workspace_bytes_scaled", "nmf_render_rays", "nmf_render_rays_host", "nmf_sample_rays",
            "nmf_vm_density", "nmf_vm_appfeature", "nmf_vm_normals", "nmf_env_lookup", "nmf_ggx_sample",
            "nmf_brdf_mlp", "nmf_material_heads", "nmf_dense_alpha", "nmf_generate_rays", "nmf_image_sq_error",
            "nmf_sample_rays_train", "nmf_train_workspace_bytes", "nmf_train_plain", "nmf_upsample_bilinear",
            "nmf_render_train_workspace_bytes", "nmf_render_rays_train", "nmf_l1_reg", "nmf_grad_sq_norm", "nmf_adam_step",
            "nmf_env_lookup_bwd_scatter", "nmf_env_lookup_bwd_finish", "nmf_env_lookup_bwd_mipbias", "nmf_vm_normals_bwd_scatter",
            "nmf_vm_normals_bwd_finish", "nmf_material_heads_bwd", "nmf_train_microfacet", "nmf_bench_gather", "nmf_pack_factor", "nmf_env_build_sat",
            "nmf_occupancy_from_alpha", "nmf_transpose_batch", "nmf_env_pair_sat", "nmf_pack_shading", "nmf_env_build_sat_dev", "nmf_env_lookup_bwd_finish_dev"]
