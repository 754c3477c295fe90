"""ctypes binding of libgraspnerf_b200.so (include/graspnerf_b200.h).

There is NO fallback: if the library is missing or a struct size disagrees the import of the product path fails.
"""
import ctypes as C
import os

from .build import LIBPATH

c_float_p = C.POINTER(C.c_float)


class GnK1Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('imgs', 'img_feats', 'ray_feats', 'KRt', 'cam', 'axis', 'bbox_min', 'pts',
                                          'que_dir', 'rec', 'pt', 'dbg_feat_idx')] + \
               [(n, C.c_int) for n in ('B', 'V', 'H', 'W', 'fh', 'fw', 'R', 'N', 'dn', 'volume_mode',
                                       'tiles_per_scene', 'feat_stride', 'img_u8')] + [('valid_count', C.c_void_p)]


class GnK2aParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('rec', 'pt', 'weights', 'depth_range', 'que_dists', 'pooled', 'colors',
                                          'dbg_rows', 'tc_const', 'tok', 'axis', 'bbox_min', 'pts')] + \
               [(n, C.c_int) for n in ('B', 'N', 'V', 'dn', 'with_rgb', 'R', 'volume_mode')] + [('status', C.c_void_p)]


class GnK2bParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('pooled', 'tok', 'weights', 'axis', 'bbox_min', 'pts', 'pos_table', 'sdf', 'grad')] + \
               [(n, C.c_int) for n in ('B', 'N', 'dn', 'R', 'volume_mode')]


class GnK3Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('sdf', 'grad', 'colors', 'que_dir', 'depth')] + \
               [('inv_s', C.c_float), ('cos_anneal_ratio', C.c_float)] + \
               [(n, C.c_void_p) for n in ('alpha', 'hit_prob', 'pixel_colors', 'render_depth', 'eik_partial')] + \
               [(n, C.c_int) for n in ('B', 'rn', 'dn')]


class GnK2bBwdParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('pooled', 'weights', 'axis', 'bbox_min', 'pts', 'pos_table', 'd_sdf', 'd_pooled', 'd_weights')] + \
               [(n, C.c_int) for n in ('B', 'N', 'dn', 'R', 'volume_mode')]


class GnK2aBwdParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('rec', 'pt', 'weights', 'depth_range', 'que_dists', 'd_pooled', 'd_rec', 'd_weights', 'd_colors')] + \
               [(n, C.c_int) for n in ('B', 'N', 'V', 'dn')]


class GnK1BwdParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('KRt', 'axis', 'bbox_min', 'pts', 'd_rec', 'd_img_feats', 'd_ray_feats')] + \
               [(n, C.c_int) for n in ('B', 'V', 'H', 'W', 'fh', 'fw', 'R', 'N', 'volume_mode')]


class GnRaySetupParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('coords', 'poses', 'Ks', 'depth', 'depth_range', 'pts', 'que_dir', 'inv_dists', 'centers', 'dirs')] + \
               [(n, C.c_int) for n in ('B', 'rn', 'dn')]


class GnDepthMeanParams(C.Structure):
    _fields_ = [('feats', C.c_void_p), ('coords', C.c_void_p), ('w_coarse', C.c_void_p * 6), ('w_fine', C.c_void_p * 6),
                ('mean', C.c_void_p), ('mean_fine', C.c_void_p)] + [(n, C.c_longlong) for n in ('stride_v', 'stride_c', 'stride_y', 'stride_x')] + \
               [(n, C.c_int) for n in ('V', 'num', 'H', 'W', 'fh', 'fw', 'align_corners')]


class GnGraspPostParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('tsdf', 'qual', 'rot', 'width', 'qual_out', 'scratch', 'grasps', 'count')] + \
               [(n, C.c_float) for n in ('sigma', 'min_width', 'max_width', 'tsdf_thres_high', 'tsdf_thres_low', 'threshold')] + \
               [(n, C.c_int) for n in ('R', 'max_filter_size', 'max_grasps')]


class GnVgnParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('volume', 'weights', 'workspace', 'out')] + [(n, C.c_int) for n in ('B', 'R', 'out_scene_stride')]


class GnNormActPadParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('x', 'gamma', 'beta', 'res', 'res_gamma', 'res_beta', 'out_padded', 'out_unpadded')] + \
               [(n, C.c_int) for n in ('N', 'C', 'H', 'W', 'pad', 'x_pad', 'res_pad', 'act')] + [('eps', C.c_float), ('x_splits', C.c_int), ('x_split_stride', C.c_longlong)]


class GnConvParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('in_', 'wimg', 'koff', 'bias', 'out')] + [('M', C.c_longlong)] + \
               [(n, C.c_int) for n in ('Nimg', 'Cin', 'Hp', 'Wp', 'Cout', 'Npad', 'Ho', 'Wo', 'stride', 'Kpad', 'ksplit')]


_lib = None


def load():
    """Loads the library (once) and declares the prototypes.  Raises if it is absent or inconsistent."""
    global _lib
    if _lib is not None:
        return _lib
    libpath = LIBPATH[:-3] + os.environ.get('GN_LIB_TAG', '') + '.so'      # GN_LIB_TAG: development A/B builds (build.py)
    if not os.path.exists(libpath):
        raise RuntimeError(
            f'{libpath} not found: build it with `python -m graspnerf_b200.build` (or __graft_entry__.build()). '
            'graspnerf_b200 has no CPU / PyTorch fallback for its CUDA kernels.')
    lib = C.CDLL(libpath)
    lib.gn_version.restype = C.c_char_p
    for name, st in (('gn_k1_forward', GnK1Params), ('gn_k2a_forward', GnK2aParams), ('gn_k2a_forward_tc', GnK2aParams), ('gn_k2b_forward', GnK2bParams),
                     ('gn_k3_composite', GnK3Params), ('gn_k2b_backward', GnK2bBwdParams), ('gn_k2a_backward', GnK2aBwdParams),
                     ('gn_k1_backward', GnK1BwdParams), ('gn_k3_ray_setup', GnRaySetupParams), ('gn_k3_depth_mean', GnDepthMeanParams), ('gn_k4_grasp_post', GnGraspPostParams), ('gn_vgn_forward', GnVgnParams), ('gn_k6_norm_act_pad', GnNormActPadParams), ('gn_k7_conv_forward', GnConvParams)):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(st), C.c_void_p]
    lib.gn_k3_coarse_depths.restype = C.c_int
    lib.gn_k3_coarse_depths.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.gn_k3_fine_depths.restype = C.c_int
    lib.gn_k3_fine_depths.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p]
    lib.gn_k2a_tc_prepare.restype = C.c_int
    lib.gn_k2a_tc_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gn_k6_upsample2x_pad.restype = C.c_int
    lib.gn_k6_upsample2x_pad.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.gn_k6_fuse_features.restype = C.c_int
    lib.gn_k6_fuse_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.gn_k6_images_u8.restype = C.c_int
    lib.gn_k6_images_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.gn_vgn_layer_info.restype = C.c_int
    lib.gn_vgn_layer_info.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 7
    lib.gn_vgn_workspace_floats.argtypes = [C.c_int]
    lib.gn_weight_entry.restype = C.c_int
    lib.gn_weight_entry.argtypes = [C.c_int, C.POINTER(C.c_char_p)] + [C.POINTER(C.c_int)] * 4
    for name, st in (('gn_sizeof_k1_params', GnK1Params), ('gn_sizeof_k2a_params', GnK2aParams),
                     ('gn_sizeof_k2b_params', GnK2bParams), ('gn_sizeof_k3_params', GnK3Params),
                     ('gn_sizeof_k2b_bwd_params', GnK2bBwdParams), ('gn_sizeof_k2a_bwd_params', GnK2aBwdParams),
                     ('gn_sizeof_k1_bwd_params', GnK1BwdParams), ('gn_sizeof_ray_setup_params', GnRaySetupParams), ('gn_sizeof_depth_mean_params', GnDepthMeanParams),
                     ('gn_sizeof_grasp_post_params', GnGraspPostParams), ('gn_sizeof_vgn_params', GnVgnParams),
                     ('gn_sizeof_norm_act_pad_params', GnNormActPadParams), ('gn_sizeof_conv_params', GnConvParams)):
        got = getattr(lib, name)()
        if got != C.sizeof(st):
            raise RuntimeError(f'{name}: library says {got} bytes, ctypes mirror has {C.sizeof(st)}')
    _lib = lib
    return lib


def weight_table():
    """[(name, offset, rows, cols, cols_padded)] as enumerated by the library."""
    lib = load()
    out = []
    for i in range(lib.gn_weight_entry_count()):
        name = C.c_char_p()
        off, rows, cols, cp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = lib.gn_weight_entry(i, C.byref(name), C.byref(off), C.byref(rows), C.byref(cols), C.byref(cp))
        assert rc == 0
        out.append((name.value.decode(), off.value, rows.value, cols.value, cp.value))
    return out


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f'{what} failed with code {rc}' + (' (argument error)' if rc < 0 else ' (cudaError_t)'))
